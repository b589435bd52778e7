"""Device time of the rate kernels alone (MODE 0: Neurons.update() without a queued motion step) next to the
skewed riab_run step, per workload -- separates the consumers' cost from the float64 motion producers'.
  python scripts/rates_only.py [c2 c2e c3 ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ratinabox_b200 as rb  # noqa: E402


def run(name, spikes, steps=200):
    wl = bench.WORKLOADS[name]
    A = wl["agents"]
    np.random.seed(0)
    Env = rb.Environment()
    for w in wl["walls"]:
        Env.add_wall(w)
    Ag = rb.Agent(Env, {"dt": 0.01, "n_agents": A, "seed": 7})
    pos, vel = bench.synthetic_agents(A, wl["walls"], 100)
    Ag.pos, Ag.velocity = pos, vel
    pops = bench.build_populations(rb, Ag, wl)
    for ns in pops:
        ns.save_spikes = spikes
    out = {}
    for mode in ("rates_only", "motion_only", "run"):
        def body(n):
            if mode == "rates_only":
                for _ in range(n):
                    for ns in pops:
                        ns.update()
            elif mode == "motion_only":
                for _ in range(n):
                    Ag.update()
                Ag._flush_pending()
            else:
                Ag.run(n)
        body(10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); body(steps); e1.record()
        torch.cuda.synchronize()
        out[mode] = round(e0.elapsed_time(e1) / steps * 1e3, 1)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or ["c2", "c2e", "c3"]
    for n in names:
        for spikes in (False, True):
            print(json.dumps({"workload": n, "spikes": spikes, "us_per_step": run(n, spikes)}), flush=True)
