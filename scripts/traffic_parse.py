"""gpurun_out/r02_traffic_<wl>.csv (ncu range replay, scripts/traffic_round.sh) -> profiles/traffic_<wl>.json."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
FLUSH = 256 << 20
for wl in ("c2", "c2e", "c3", "c4"):
    path = os.path.join(ROOT, "gpurun_out", f"r02_traffic_{wl}.csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    col = {k: i for i, k in enumerate(hdr)}
    tot = {}
    for r in rows:
        if r is hdr or len(r) < len(hdr):
            continue
        name, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], r[col["Metric Value"]].replace(",", "")
        if name.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            tot[name] = tot.get(name, 0.0) + float(val) * mult
    if not tot:
        continue
    w = bench.WORKLOADS[wl]
    n_cells = sum(n for _, n, _ in bench.cells_of(w))
    algo = bench.algorithmic_bytes_per_agent_step(n_cells, True) * w["agents"]
    per_step = (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"] - FLUSH) / steps
    out = {"dram_bytes_per_step": per_step, "algorithmic_bytes_per_step": algo, "ratio": per_step / algo,
           "read_bytes_range": tot["dram__bytes_read.sum"], "write_bytes_range": tot["dram__bytes_write.sum"], "steps_in_range": steps,
           "flush_bytes_subtracted": FLUSH,
           "source": f"ncu --replay-mode range over {steps} consecutive riab_run steps + one {FLUSH >> 20} MiB flush write "
                     f"(scripts/traffic_round.sh, scripts/prof_driver.py mode 'range')"}
    json.dump(out, open(os.path.join(ROOT, "profiles", f"traffic_{wl}.json"), "w"), indent=1)
    print(wl, f"{per_step / 1e6:.1f} MB/step vs algorithmic {algo / 1e6:.1f} MB  ratio {per_step / algo:.3f}")
