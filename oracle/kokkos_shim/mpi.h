// mpi.h -- stub so that /root/reference/src/lpm_comm.hpp parses (oracle/_ref only; nothing MPI runs).
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return 0; }
inline int MPI_Initialized(int* f) { *f = 1; return 0; }
inline double MPI_Wtime() { return 0.0; }
#endif
