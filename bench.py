#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the RatInABox per-step hot path on B200.

One "step" = Agent.update() + Neurons.update() of every population for every agent
(BASELINE.json metric).  Default workload = BASELINE.json configs[1]:
65 536 agents, 1x1 m box + 2 internal walls, 1 024 Gaussian PlaceCells with the
reference's default wall geometry for that box (geodesic -> line_of_sight,
ratinabox/Neurons.py:922-928), dt = 10 ms, history + spikes on (reference defaults).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2e|c3|c4]

N > 1 is launched by torchrun (one rank per GPU); agents are sharded (weak scaling:
65 536 agents per GPU), there is no collective on the step path.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOX_WALLS = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]


def maze_walls(n=8, length=0.6):
    out = []
    for k in range(1, n + 1):
        x = k / (n + 1)
        out.append([[x, 0.0], [x, length]] if k % 2 else [[x, 1.0], [x, 1.0 - length]])
    return out


WORKLOADS = {
    # name: (agents per GPU, walls, population spec, description)
    "c2": dict(agents=65536, walls=BOX_WALLS, cells=("place", 1024, "line_of_sight"),
               desc="configs[1]: 65536 agents, box+2 walls, 1024 Gaussian PlaceCells (reference-default line_of_sight), dt=10ms"),
    "c2e": dict(agents=65536, walls=BOX_WALLS, cells=("place", 1024, "euclidean"),
                desc="configs[1] with wall_geometry='euclidean'"),
    "c3": dict(agents=65536, walls=[], cells=("grid", 1024, None),
               desc="configs[2]: 65536 agents, 1024 GridCells, box"),
    "c4": dict(agents=16384, walls=maze_walls(), cells=("bvc", 512, None),
               desc="configs[3]: 16384 agents, 512 BVCs, 8-wall maze (4 boundary + 8 internal)"),
    # configs[4] is a STRONG-scaling case: 262144 agents in total, divided over the ranks
    "c5": dict(agents=262144, strong=True, walls=BOX_WALLS,
               cells=[("place", 512, "line_of_sight"), ("grid", 512, None), ("bvc", 256, None)],
               desc="configs[4]: 262144 agents in total, 512 Place + 512 Grid + 256 BVC populations, box+2 walls"),
}


def cells_of(wl):
    c = wl["cells"]
    return list(c) if isinstance(c, list) else [c]


def agents_per_rank(wl, world):
    return wl["agents"] // world if wl.get("strong") else wl["agents"]


def synthetic_cells(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "place":
        g = int(round(np.sqrt(n)))
        delta = 1.0 / g
        xs = np.linspace(delta / 2, 1 - delta / 2, g)
        c = np.array(np.meshgrid(xs, xs)).reshape(2, -1).T[:n]
        c = c + rng.uniform(-0.45 * delta, 0.45 * delta, c.shape)      # uniform_jitter (Environment.py:601-631)
        return dict(centres=c, widths=0.2 * np.ones(len(c)))
    if kind == "grid":
        return dict(gridscales=rng.uniform(0.2, 1.0, n), orientations=rng.uniform(0, np.pi / 3, n),
                    phase_offsets=rng.uniform(0, 2 * np.pi, (n, 2)))
    if kind == "bvc":
        mu_d = rng.uniform(0.05, 0.3, n)
        return dict(mu_d=mu_d, sg_d=0.08 + mu_d / 12, mu_t=rng.uniform(0, 360, n), sg_t=rng.uniform(10, 30, n))
    raise ValueError(kind)


def synthetic_agents(n, walls, seed):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.02, 0.98, (n, 2))
    ang = rng.uniform(0, 2 * np.pi, n)
    vel = 0.08 * np.stack((np.cos(ang), np.sin(ang)), axis=1)
    return pos, vel


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions: one streaming
    `nvidia-smi -lms 50` process (per-call start-up would miss short regions)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._p, self._t = index, [], None, None

    def _pump(self):
        try:
            for line in self._p.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
        except Exception:
            pass

    def start(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                        "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._pump, daemon=True)
            self._t.start()
            time.sleep(0.3)            # let the first samples arrive before the timed region starts
            self.rows.clear()
        except Exception:
            self._p = None

    def stop(self):
        if self._p is not None:
            time.sleep(0.06)
            self._p.terminate()
            try:
                self._p.wait(timeout=5)
            except Exception:
                self._p.kill()
            if self._t:
                self._t.join(timeout=2)
        num = lambda x: x.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in self.rows if num(r[0])]
        mx = [float(r[1]) for r in self.rows if num(r[1])]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------- CPU baseline (port)
def _oracle_worker(args):
    """One process = one reference-style Agent + population stepped in a Python loop
    (how the reference runs: SURVEY.md section 3.1).  Returns (agent_steps, seconds)."""
    wl_name, n_steps, seed = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import riab_oracle as O
    wl = WORKLOADS[wl_name]
    np.random.seed(seed)
    env = O.OracleEnvironment(walls=wl["walls"])
    pos, vel = synthetic_agents(1, wl["walls"], seed)
    ag = O.OracleAgent(env, pos[0], vel[0], {"dt": 0.01})
    rng = O.GlobalRNG()

    def population(kind, n, geom, k):
        cp = synthetic_cells(kind, n, k)
        if kind == "place":
            return O.OracleNeurons(ag, n, lambda p, r: O.place_cells_get_state(env, cp["centres"], cp["widths"], p, r,
                                                                               "gaussian", geom))
        if kind == "grid":
            w = O.grid_cells_w(cp["orientations"])
            return O.OracleNeurons(ag, n, lambda p, r: O.grid_cells_get_state(cp["gridscales"], cp["phase_offsets"], w, p))
        return O.OracleNeurons(ag, n, lambda p, r: O.bvc_get_state(env, cp["mu_d"], np.radians(cp["mu_t"]), cp["sg_d"],
                                                                   np.radians(cp["sg_t"]), p, r))

    pops = [population(kind, n, geom, k) for k, (kind, n, geom) in enumerate(cells_of(wl))]
    for _ in range(5):
        ag.update(rng)
        for ns in pops:
            ns.update(rng)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        ag.update(rng)
        for ns in pops:
            ns.update(rng)
    return n_steps, time.perf_counter() - t0


def cpu_port_rate(wl_name, n_steps, procs):
    """agent-steps/s of the NumPy port on `procs` host processes (independent agents)."""
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_worker((wl_name, n_steps, 0))]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_oracle_worker, [(wl_name, n_steps, s) for s in range(procs)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    inner = max(r[1] for r in res)
    return steps / inner, wall


# -------------------------------------------------------------------------------- main
def build_populations(rb, Ag, wl):
    return [build_population(rb, Ag, kind, n, geom, k) for k, (kind, n, geom) in enumerate(cells_of(wl))]


def build_population(rb, Ag, kind, n, geom, k=0):
    cp = synthetic_cells(kind, n, k)
    if kind == "place":
        return rb.PlaceCells(Ag, {"place_cell_centres": cp["centres"], "widths": 0.2, "description": "gaussian",
                                  "wall_geometry": geom})
    if kind == "grid":
        return rb.GridCells(Ag, {"gridscale": cp["gridscales"], "orientation": cp["orientations"],
                                 "phase_offset": cp["phase_offsets"]})
    # The reference's VectorCells declare `angular_spread` in default_params but hand **params to
    # utils.create_random_assembly, which reads `sigma_angle` (ratinabox/Neurons.py:1315 vs utils.py:1124-1131): a
    # `sigma_angle` key works but warns, an `angular_spread` key is silently ignored.  The tuning is therefore set the way
    # the reference tells its users to (Neurons.py:1612): by assigning the arrays after construction.
    B = rb.BoundaryVectorCells(Ag, {"n": n})
    B.tuning_distances, B.tuning_angles = np.asarray(cp["mu_d"], float), np.radians(cp["mu_t"])
    B.sigma_distances, B.sigma_angles = np.asarray(cp["sg_d"], float), np.radians(cp["sg_t"])
    return B


def algorithmic_bytes_per_agent_step(n_cells, spikes):
    """DESIGN.md 'bytes per unit': float64 state in+out (12 doubles each), float32 agent
    history row (8), one float32 write per rate, bit-packed spikes."""
    return 2 * 12 * 8 + 8 * 4 + 4 * n_cells + (n_cells // 8 if spikes else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spikes", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cl = cells_of(wl)
    n_cells = sum(n for _, n, _ in cl)
    kind = "+".join(k for k, _, _ in cl)
    geom = cl[0][2]
    A_rank = agents_per_rank(wl, world)
    scaling = "strong" if wl.get("strong") else "weak"
    config = {"workload": f"{args.workload}: {wl['desc']}", "agents_per_gpu": A_rank, "n_cells": n_cells,
              "cells": kind, "wall_geometry": geom, "n_walls": 4 + len(wl["walls"]), "dt": 0.01,
              "spikes": not args.no_spikes, "history": "device rings (rates: last rows within 8 GiB; agent rows: all)",
              "l2": "each step writes >= 2x L2 of fresh rate rows (inputs larger than L2)",
              "parallelism": f"agents sharded x{world}, no step-path collective"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        # one single-threaded process per host core (the reference is single-threaded; BLAS/OpenMP pools
        # in 128 workers would only oversubscribe the box)
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
            os.environ[v] = "1"
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        per = {"c2": 300, "c2e": 400, "c3": 400, "c4": 60, "c5": 60}[args.workload]
        vals = []
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_port_rate(args.workload, max(10, per // 10), cores)
        t_all = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            v, _ = cpu_port_rate(args.workload, per, cores)
            vals.append(v)
        value = float(np.mean(vals))
        line = {"impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * A_rank * world / value, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                 "sample": f"{cores} processes x 1 agent x {per} steps of the NumPy port (oracle/riab_oracle.py), "
                                           f"{len(vals)} repeats; /root/reference is Python and cannot travel to the GPU box"},
                "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
        print(json.dumps(line))
        return

    # ----------------------------------------------------------------------- our arm
    import torch
    import ratinabox_b200 as rb
    from ratinabox_b200 import _lib
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    A = A_rank
    np.random.seed(1234 + rank)
    Env = rb.Environment()
    for w in wl["walls"]:
        Env.add_wall(w)
    Ag = rb.Agent(Env, {"dt": 0.01, "n_agents": A, "seed": 7, "id_offset": rank * A})
    pos, vel = synthetic_agents(A, wl["walls"], 100 + rank)
    Ag.pos, Ag.velocity = pos, vel
    Ag.measured_velocity = vel
    pops = build_populations(rb, Ag, wl)
    if args.no_spikes:
        for ns in pops:
            ns.save_spikes = False
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): riab_run, no host work between steps
    Ag.run(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.riab_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    Ag.run(args.steps)
    ev1.record()
    barrier()
    launches = lib.riab_launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * A * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps          # one fused kernel per step (BVC: two) -> per-step device time

    # ---- e2e: the Python API with HOST buffers each step (drift in, positions out)
    e2e_steps = max(10, min(args.steps, 500))
    drift = (0.05 * torch.randn((A, 2), dtype=torch.float64)).pin_memory()       # a policy's velocity commands
    for _ in range(3):
        Ag.update(drift_velocity=drift)
        for ns in pops:
            ns.update()
        _ = Ag.state_view("pos")
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Ag.update(drift_velocity=drift)          # H2D: A*2*8 B read from pinned host memory by the motion kernel, every step
        for ns in pops:
            ns.update()                          # first population: fused motion + rates kernel; others: rates
        p = Ag.state_view("pos")                 # D2H: A*2*8 B posted to pinned host memory by the motion kernel; blocks until
                                                 # the whole step (motion + rates) has finished (`Ag.pos` = a private copy of it)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()        # sampled across the device-resident and the e2e timed regions
    e2e = {"value": world * A * e2e_steps / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": A * 16,
           "d2h_bytes_per_step": A * 16, "steps": e2e_steps,
           "api": "Agent.update(drift_velocity=<pinned host tensor>) + Neurons.update() + Agent.state_view('pos') (host positions), per step; rates stay in the device history ring (268 MB/step at c2 cannot cross PCIe)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    spikes = not args.no_spikes
    bytes_unit = algorithmic_bytes_per_agent_step(n_cells, spikes)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_unit * A / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                "bytes_per_agent_step": bytes_unit, "kernel_ms": kernel_ms}
    n_bvc = sum(n for k, n, _ in cl if k == "bvc")
    if n_bvc:
        # BoundaryVectorCells are bound by the special-function unit, not by HBM: one ex2 per (agent, cell, test angle)
        # in the angular integral (T = 180) against 16 MUFU results per clock per SM (148 SMs at the sampled SM clock)
        ex2 = A * n_bvc * 180 / (kernel_ms * 1e-3)
        peak_ex2 = 148 * 16 * (clocks.get("sm_mhz") or 1965.0) * 1e6
        roofline["mufu"] = {"achieved_ex2_per_s": ex2, "peak_ex2_per_s": peak_ex2, "frac": ex2 / peak_ex2,
                            "note": "share of the whole step (motion, rays, other populations included) spent at the ex2 rate"}
    prof = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:          # a reported baseline, timed at N=1 only
        per = {"c2": 4000, "c2e": 5000, "c3": 5000, "c4": 800, "c5": 600}[args.workload]
        v, wall = cpu_port_rate(args.workload, per, 1)
        cpu_baseline = {"value": v, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                        "sample": f"1 agent x {per} steps of oracle/riab_oracle.py (NumPy port, same cost structure as the "
                                  f"reference's per-agent Python loop), {wall:.1f} s"}
    line = {"metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32 rates / f64 agent state", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
