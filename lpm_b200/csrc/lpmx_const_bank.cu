// lpmx_const_bank.cu -- one constant bank of the constant-bank velocity path: compiled once per bank with -DLPMX_CS_BANK=<n>
// (lpm_b200/build.py), each object its own module = its own 64 KB user constant bank (lpmx_const_bank.cuh, lpmx_const_stream.cu).
#include "lpmx_const_bank.cuh"
