#!/bin/bash
# r3g (1 GPU): the GMLS pivot-ratio check on the device: every test that goes through lpmx_gmls.cu (Laplacian, interpolation /
# remesh, AMR drivers, examples)
TAG=${1:-r3g}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gmls.py tests/test_amr.py tests/test_examples.py tests/test_gpu_parity_swe_rk2.py -q -m gpu -x 2>&1 | tail -4 | tee $OUT/pytest_gmls.log
