"""Time the stand-alone Agent.update kernel (riab_agent_update) and a rates-only pass."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ratinabox_b200 as rb
A = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
E = rb.Environment()
for w in [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]:
    E.add_wall(w)
Ag = rb.Agent(E, {"dt": 0.01, "n_agents": A, "save_history": False})
for _ in range(5):
    Ag.update()
Ag._flush_pending(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    Ag.update()          # each update flushes the previous queued step through k_agent_update
Ag._flush_pending()
e1.record(); torch.cuda.synchronize()
res = {"agents": A, "motion_us_per_step": e0.elapsed_time(e1) / 50 * 1e3}
PCs = rb.PlaceCells(Ag, {"n": 1024, "wall_geometry": "euclidean", "save_spikes": False})
for _ in range(3):
    PCs.update()
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    PCs.update()         # no queued motion -> rates-only kernel (producers only build records)
e1.record(); torch.cuda.synchronize()
res["rates_only_us_per_step"] = e0.elapsed_time(e1) / 50 * 1e3
print(json.dumps(res))
