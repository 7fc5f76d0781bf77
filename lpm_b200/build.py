"""In-tree build of liblpmx.so (sm_100a CUDA kernels + C ABI + host mesh generator).

Usage:  python -m lpm_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the resulting lpm_b200/liblpmx.so is git-ignored
but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "liblpmx.so")

CU_SOURCES = ["lpmx_core.cu", "lpmx_peer.cu", "lpmx_kernels.cu", "lpmx_const_stream.cu", "lpmx_sums.cu", "lpmx_steppers.cu", "lpmx_swe_stepper.cu", "lpmx_plane.cu", "lpmx_diagnostics.cu", "lpmx_gmls.cu", "lpmx_refinement.cu"]
# lpmx_const_bank.cu is compiled once per constant bank: every object is its own module with its own 64 KB bank
# (kCsBanks in csrc/lpmx_internal.h must equal N_CONST_BANKS)
N_CONST_BANKS = 24
CXX_SOURCES = ["lpmx_mesh.cpp"]
HEADERS = ["lpmx_internal.h", "lpmx_finalize.cuh", "lpmx_pair_kernel.cuh", "lpmx_gmls_core.h", "lpmx_fast_log.h", "lpmx_peer_protocol.h", "lpmx_const_stream_body.h", "lpmx_const_bank.cuh", "seed_tables.inc", "log_table_7.inc", "log_table_8.inc", "log_table_10.inc", os.path.join("..", "..", "include", "lpmx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]
# the mesh generator is compiled without FMA contraction so coordinates do not depend on the compiler
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose, log):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose:
        sys.stdout.write(p.stdout + p.stderr)


def build(force=False, verbose=False, jobs=None):
    """Compile what is stale (all of it with force) and link; the compile steps are independent and run `jobs` at a time
    (default: the host's cores, at most 8)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    cmds = []
    for src in CU_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src + ".o")
        if force or _stale(obj, [path] + hdrs):
            cmds.append([_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj])
        objs.append(obj)
    bank_src = os.path.join(CSRC, "lpmx_const_bank.cu")
    for k in range(N_CONST_BANKS):
        obj = os.path.join(OBJ, "lpmx_const_bank%d.cu.o" % k)
        if force or _stale(obj, [bank_src] + hdrs):
            cmds.append([_nvcc()] + NVCC_FLAGS + ["-DLPMX_CS_BANK=%d" % k, "-c", bank_src, "-o", obj])
        objs.append(obj)
    for src in CXX_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src + ".o")
        if force or _stale(obj, [path] + hdrs):
            cmds.append(["g++"] + CXX_FLAGS + ["-c", path, "-o", obj])
        objs.append(obj)
    log = []
    if cmds:
        jobs = jobs or max(1, min(8, os.cpu_count() or 1))
        logs = [[] for _ in cmds]
        with ThreadPoolExecutor(max_workers=jobs) as pool:
            for f in [pool.submit(_run, c, verbose, lg) for c, lg in zip(cmds, logs)]:
                f.result()  # re-raises the first failure
        for lg in logs:
            log += lg
    if force or _stale(LIB, objs):
        _run([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"], verbose, log)
    with open(os.path.join(OBJ, "build.log"), "a") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
