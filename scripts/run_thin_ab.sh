#!/bin/bash
# A/B of the spike streams on one B200: GPU tests, then c2 / c2e / c3 with the thinned (default) and the dense stream.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_gpu_tests.txt
B="python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline"
$B > gpurun_out/r02b_c2_thin.json 2> gpurun_out/r02b_c2_thin.err
RIAB_DENSE_SPIKES=1 $B > gpurun_out/r02b_c2_dense.json 2>/dev/null
$B --no-spikes > gpurun_out/r02b_c2_nospikes.json 2>/dev/null
for w in c2e c3; do
  $B --workload $w > gpurun_out/r02b_${w}_thin.json 2>/dev/null
  RIAB_DENSE_SPIKES=1 $B --workload $w > gpurun_out/r02b_${w}_dense.json 2>/dev/null
done
python scripts/rates_only.py c2 c2e c3 2>/dev/null | grep -v Warn > gpurun_out/r02b_rates_only.txt
tail -3 gpurun_out/r02b_gpu_tests.txt
for f in gpurun_out/r02b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"]*1e3,1), "us", "%.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
cat gpurun_out/r02b_rates_only.txt
