#!/bin/bash
# DRAM bytes per step from an ncu RANGE over N consecutive steps of riab_run (see scripts/prof_driver.py, mode "range").
#   scripts/traffic_round.sh [steps]   -> gpurun_out/r02_traffic_<workload>.csv ; scripts/traffic_parse.py turns them into
#   profiles/traffic_<workload>.json (dram_bytes_per_step = (read + write - flush bytes) / steps)
steps=${1:-6}
mkdir -p gpurun_out
for w in c2 c2e c3 c4; do
  timeout 300 ncu --replay-mode range --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      --csv --log-file gpurun_out/r02_traffic_$w.csv python scripts/prof_driver.py $w 1 range $steps > gpurun_out/r02_traffic_$w.log 2>&1
  tail -3 gpurun_out/r02_traffic_$w.csv
done
