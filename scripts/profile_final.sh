#!/bin/bash
# Round-2 closing pass on one B200 (gpurun), after the BVC ray-screen rewrite: GPU tests, BVC bench lines, launch list,
# full ncu captures of the two BVC kernels, the driver-style default bench line.  Output: gpurun_out/ (CSV / JSON only).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_gpu_tests.log
python bench.py --workload c4 --steps 100 --warmup 10 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2>/dev/null
RIAB_BENCH_BVC_SIGMA_DEG=11.25 python bench.py --workload c4 --steps 100 --warmup 10 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c4_sigma11.json 2>/dev/null
python bench.py --workload c5 --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c5.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 60 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --workload c4 --steps 8 --warmup 2 --no-extra --no-cpu-baseline > /dev/null 2>&1
scripts/ncu_export.sh r02_c4_integrate k_bvc_integrate 2 c4 1 run 4 > /dev/null
scripts/ncu_export.sh r02_c4_rays k_bvc_rays 2 c4 1 run 4 > /dev/null
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
ls -la gpurun_out | tail -20
