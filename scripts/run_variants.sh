#!/bin/bash
# A/B of library variants (ratinabox_b200/variants/lib<name>.so, built with extra -D flags; "intree" = the shipped library)
# on one B200; every process under `timeout` (a variant whose setmaxnreg split does not fit hangs).
#   scripts/run_variants.sh <tag> <name> [<name> ...]
set -u
tag=$1; shift
mkdir -p gpurun_out
V=ratinabox_b200/variants
B="timeout 120 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline"
for name in "$@"; do
  lib=$V/lib$name.so
  [ "$name" = intree ] && lib=ratinabox_b200/libriab_b200.so
  RIAB_LIB=$lib $B --workload c2 > gpurun_out/${tag}_${name}_c2.json 2> gpurun_out/${tag}_${name}_c2.err || { echo "$name: c2 failed / timed out"; continue; }
  RIAB_LIB=$lib $B --workload c2 --no-spikes > gpurun_out/${tag}_${name}_c2_nospikes.json 2>/dev/null
  for w in c2e c3; do
    RIAB_LIB=$lib $B --workload $w > gpurun_out/${tag}_${name}_${w}.json 2>/dev/null
  done
  RIAB_LIB=$lib timeout 120 python scripts/rates_only.py c2 2>/dev/null | grep -v Warn > gpurun_out/${tag}_${name}_rates_only.txt
done
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"]*1e3,1), "us", "%.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
cat gpurun_out/${tag}_*_rates_only.txt
