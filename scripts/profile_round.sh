#!/bin/bash
# Round-2 measurement pass on one B200 (gpurun): bench lines, launch lists, full ncu captures, range-replay DRAM traffic.
# Everything lands in gpurun_out/ (CSV / JSON only: an .ncu-rep with sources is ~75 MB).
set -u
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
python bench.py --steps 200 --warmup 20 --no-spikes --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c2_nospikes.json 2>/dev/null
for w in c2e c3; do
  python bench.py --workload $w --steps 200 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2>/dev/null
  python bench.py --workload $w --steps 200 --warmup 20 --no-extra --no-cpu-baseline --no-spikes > gpurun_out/r02_bench_${w}_nospikes.json 2>/dev/null
done
python bench.py --workload c4 --steps 100 --warmup 10 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2>/dev/null
python bench.py --workload c5 --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_c5.json 2>/dev/null
python scripts/rates_only.py c2 c2e c3 c4 2>/dev/null | grep -v Warn > gpurun_out/r02_rates_only.txt
python scripts/e2e_breakdown.py > gpurun_out/r02_e2e_breakdown.txt 2>&1
# launch lists (every launch with its device time; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 8 --warmup 2 --no-extra --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 60 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --workload c4 --steps 8 --warmup 2 --no-extra --no-cpu-baseline > /dev/null 2>&1
RIAB_NO_WHOLE_RUN=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/r02_launches_c2_perstep.csv python bench.py --steps 8 --warmup 2 --no-extra --no-cpu-baseline > /dev/null 2>&1
# full captures: the whole-run kernel (8 steps in one launch), the per-step skewed kernel, the BVC kernels
scripts/ncu_export.sh r02_c2_wholerun k_step 1 c2 1 run 8 > /dev/null
RIAB_NO_WHOLE_RUN=1 scripts/ncu_export.sh r02_c2_perstep k_step 4 c2 1 run 8 > /dev/null
scripts/ncu_export.sh r02_c4_integrate k_bvc_integrate 2 c4 1 run 4 > /dev/null
scripts/ncu_export.sh r02_c4_rays k_bvc_rays 2 c4 1 run 4 > /dev/null
rm -f gpurun_out/*.source.csv.gz.tmp
scripts/traffic_round.sh 6 > gpurun_out/r02_traffic.log 2>&1
ls -la gpurun_out | tail -40
