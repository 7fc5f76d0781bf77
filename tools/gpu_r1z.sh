#!/bin/bash
# 1-GPU visit r1z: BASELINE configs[3] at its stated size (sphere_swe_tc2, icos depth 8, SWERK2) and an ncu --set full capture
# of the spherical SWE pair-sum kernel (the dominant kernel of that config).
TAG=${1:-r1z}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== ncu full swe"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_sum -s 2 -c 1 -o $OUT/pair_sum_swe python bench.py --stepper swe_rk2 --workload tc2_cubed7 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_swe.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_full_swe.log
ncu -i $OUT/pair_sum_swe.ncu-rep --page raw --csv > $OUT/pair_sum_swe_raw.csv 2>/dev/null; ls -la $OUT
echo "== bench swe icos8"; timeout 420 python bench.py --stepper swe_rk2 --workload tc2_icos8 --laplacian gmls --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_swe_icos8.json 2> $OUT/bench_swe_icos8.err; echo "rc=$?"; cut -c1-1500 $OUT/bench_swe_icos8.json; tail -3 $OUT/bench_swe_icos8.err
