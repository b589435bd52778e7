#!/bin/bash
# Final measurement pass of round 2 on one B200 (gpurun): GPU tests, the driver-style bench line, per-workload lines,
# launch list + full ncu capture of the thinned-spike GridCells kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_final_gpu_tests.txt
tail -2 gpurun_out/r02_final_gpu_tests.txt
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_final_bench_full.json 2> gpurun_out/r02_final_bench_full.err
B="timeout 120 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline"
$B --no-spikes > gpurun_out/r02_final_c2_nospikes.json 2>/dev/null
$B --workload c3 --no-spikes > gpurun_out/r02_final_c3_nospikes.json 2>/dev/null
$B --workload c2e --no-spikes > gpurun_out/r02_final_c2e_nospikes.json 2>/dev/null
RIAB_DENSE_SPIKES=1 $B --workload c3 > gpurun_out/r02_final_c3_dense.json 2>/dev/null
timeout 120 python scripts/rates_only.py c2 c3 2>/dev/null | grep -v Warn > gpurun_out/r02_final_rates_only.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_final_launches_c3.csv python bench.py --workload c3 --steps 8 --warmup 2 --no-extra --no-cpu-baseline > /dev/null 2>&1
timeout 300 scripts/ncu_export.sh r02_final_c3_thin_run k_step 0 c3 1 run 8 > /dev/null
python scripts/ncu_source_top.py gpurun_out/r02_final_c3_thin_run.source.csv.gz 60 > gpurun_out/r02_final_c3_thin_run.top.txt 2>&1
rm -f gpurun_out/r02_final_c3_thin_run.source.csv.gz
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3,1), "us", "%.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
        for k,w in d.get("workloads",{}).items(): print("   ",k, round(w["ms_per_step"]*1e3,1), "e2e %.3g"%w["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
PY
cat gpurun_out/r02_final_rates_only.txt; head -3 gpurun_out/r02_final_c3_thin_run.top.txt
