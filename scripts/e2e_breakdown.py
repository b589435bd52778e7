"""Where the e2e step of bench.py goes: device time of the fused step vs host time of the three API calls."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
import ratinabox_b200 as rb

wl = bench.WORKLOADS["c2"]
A = wl["agents"]
np.random.seed(1)
Env = rb.Environment()
for w in wl["walls"]:
    Env.add_wall(w)
Ag = rb.Agent(Env, {"dt": 0.01, "n_agents": A, "seed": 7})
pops = bench.build_populations(rb, Ag, wl)
drift = (0.05 * torch.randn((A, 2), dtype=torch.float64)).pin_memory()
for _ in range(20):
    Ag.update(drift_velocity=drift); pops[0].update(); _ = Ag.state_view('pos')
torch.cuda.synchronize()
N = 300
t_upd = t_ns = t_pos = 0.0
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
t0 = time.perf_counter()
for i in range(N):
    a = time.perf_counter()
    Ag.update(drift_velocity=drift)
    b = time.perf_counter()
    ev[i][0].record()
    pops[0].update()
    ev[i][1].record()
    c = time.perf_counter()
    p = Ag.state_view('pos')
    d = time.perf_counter()
    t_upd += b - a; t_ns += c - b; t_pos += d - c
torch.cuda.synchronize()
tot = time.perf_counter() - t0
dev = sum(e0.elapsed_time(e1) for e0, e1 in ev) / N * 1e3
print(f"per step: total {tot/N*1e6:.1f} us | Agent.update {t_upd/N*1e6:.1f} | Neurons.update (launch) {t_ns/N*1e6:.1f} | Ag.state_view('pos') (wait) {t_pos/N*1e6:.1f} | device (events around Neurons.update) {dev:.1f} us")
