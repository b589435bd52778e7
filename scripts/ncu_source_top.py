"""Top SASS instructions by warp-stall samples from an `ncu --page source --csv` export (gz ok).
  python scripts/ncu_source_top.py <source.csv.gz> [N]     -> per-instruction rows + stall-kind totals + executed counts"""
import csv
import gzip
import sys

path = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][0], 16)
kinds = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[col["# Samples"]]) for r in data)
texec = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"{len(data)} instructions, {tot} samples, {texec} warp-instructions executed")
agg = {k: sum(int(r[col[k]] or 0) for r in data) for k in kinds}
print("stall kinds:", ", ".join(f"{k[6:]} {v} ({100*v/tot:.1f}%)" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
top = sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:N]
for r in sorted(top, key=lambda r: int(r[0], 16)):
    ks = sorted(((int(r[col[k]] or 0), k[6:]) for k in kinds), reverse=True)[:3]
    print(f"{int(r[0],16)-base:06x} {int(r[col['# Samples']]):6d} {100*int(r[col['# Samples']])/tot:5.1f}% exec {int(r[col['Instructions Executed']]):8d}  {r[1].strip():70s} " + " ".join(f"{k}:{v}" for v, k in ks if v))
