"""Tiny driver for ncu: python scripts/prof_driver.py <workload> <spikes 0|1> <mode rates|run> [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ratinabox_b200 as rb  # noqa: E402

name, spikes, mode = sys.argv[1], sys.argv[2] == "1", sys.argv[3]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
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
if mode == "rates":
    for _ in range(steps):
        for ns in pops:
            ns.update()
elif mode == "range":
    # DRAM traffic of `steps` CONSECUTIVE steps as one profiled range (ncu --replay-mode range): a single-launch capture
    # misses the dirty tail the 126 MB L2 still holds when the kernel ends; over a range of steps every step's rows are
    # evicted by the next ones (and the range ends with an L2-sized flush write so the last step's tail is counted too)
    Ag.run(8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    Ag.run(steps)
    flush.fill_(1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    Ag.run(steps)
torch.cuda.synchronize()
