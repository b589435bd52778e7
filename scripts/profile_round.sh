#!/bin/bash
# Round profile pass, run on the GPU box:   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r01'
# Writes into gpurun_out/: launch lists (ncu, per-launch device time), one `--set full` capture of the headline
# kernel, and the bench lines of every workload.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# 1. launch lists
timeout 300 $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_c2.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 140 --csv --log-file $OUT/${TAG}_launches_c4.csv \
    python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_c4.log 2>&1
# 2. the headline kernel, full set, one launch well inside the timed region
timeout 600 $NCU --set full --import-source on -k regex:k_step -s 12 -c 1 -f -o $OUT/${TAG}_k_step_c2 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
# the report can exceed what gpurun_out/ carries back (64 MiB): keep its raw and source pages as gzipped CSV
ncu -i $OUT/${TAG}_k_step_c2.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/${TAG}_k_step_c2.raw.csv.gz
ncu -i $OUT/${TAG}_k_step_c2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/${TAG}_k_step_c2.source.csv.gz
if [ $(stat -c %s $OUT/${TAG}_k_step_c2.ncu-rep) -gt 40000000 ]; then rm -f $OUT/${TAG}_k_step_c2.ncu-rep; fi
# 3. bench lines (not under ncu)
for W in c2 c2e c3; do
  python bench.py --workload $W --steps 600 --warmup 50 2> $OUT/${TAG}_bench_${W}.err | tail -1 > $OUT/${TAG}_bench_${W}.json
  python bench.py --workload $W --steps 600 --warmup 50 --no-spikes --no-cpu-baseline 2> /dev/null | tail -1 > $OUT/${TAG}_bench_${W}_nospikes.json
done
python bench.py --workload c4 --steps 300 --warmup 20 2> $OUT/${TAG}_bench_c4.err | tail -1 > $OUT/${TAG}_bench_c4.json
python bench.py --workload c5 --steps 60 --warmup 5 --no-cpu-baseline 2> $OUT/${TAG}_bench_c5.err | tail -1 > $OUT/${TAG}_bench_c5.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> /dev/null | tail -1 > $OUT/${TAG}_bench_reference.json
ls -la $OUT | tail -30
