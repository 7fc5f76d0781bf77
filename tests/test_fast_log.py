"""CPU: the pair kernels' table-driven log (lpm_b200/csrc/lpmx_fast_log.h, __host__ __device__) compiled for the host and
compared with logl over 4 million arguments spanning everything a mesh can produce (1e-16 .. 4, every table boundary and its
neighbours, powers of two, 1 +- ulp), plus the special cases (log 0 non-finite, log of a negative number = NaN, like std::log),
for the adopted table shape and for the measured alternatives kept behind the header's macros.  The GPU test tests/test_gpu_parity_bve.py::test_fast_log_accuracy_over_the_kernel_range
checks the device build of the same source against numpy."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("flags", [[], ["-DLPMX_LOG_MBITS=7"], ["-DLPMX_LOG_MBITS=10", "-DLPMX_LOG_KTAB=1"]])
def test_fast_log_host_build_accuracy(tmp_path, flags):
    exe = str(tmp_path / "fast_log_check")
    # -ffp-contract=off: the header spells out every fma; nothing else may be fused (nvcc does not contract across calls either)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", *flags, "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "fast_log_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    n, worst, worst_d = int(out[0]), float(out[1]), float(out[2])
    assert n > 4_000_000
    assert worst < 3.0e-16, (worst, worst_d)  # |fast_log - log| <= 3e-16 max(1, |log d|)
    assert out[3:] == ["1", "1"]  # log(0) non-finite; log(-1) = NaN
