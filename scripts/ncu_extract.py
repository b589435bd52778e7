"""Condense one `ncu --set full` report into the JSON committed under profiles/.

  python scripts/ncu_extract.py gpurun_out/r01_k_step_c2.{ncu-rep|raw.csv.gz} profiles/r01_ncu_k_step_c2.json [profiles/traffic_c2.json]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv.gz"):          # raw page exported on the GPU box (scripts/profile_round.sh)
        import gzip
        txt = gzip.open(rep, "rt").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {"kernel": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "metrics": {}, "stalls_per_issue": {}}
    for h, u, v in zip(hdr, units, val):
        if h in KEEP:
            d["metrics"][h] = {"value": float(v.replace(",", "")), "unit": u}
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            d["stalls_per_issue"][h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
    json.dump(d, open(out, "w"), indent=1)
    m = d["metrics"]
    if len(sys.argv) > 3:
        def b(k):
            x = m[k]
            return x["value"] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[x["unit"]]
        tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        json.dump({"dram_bytes_per_launch": tot, "source": rep.split("/")[-1] + " (ncu --set full, one launch)"},
                  open(sys.argv[3], "w"), indent=1)
    print(json.dumps(d["metrics"], indent=1)[:1500])


if __name__ == "__main__":
    main()
