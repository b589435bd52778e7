"""The reference-side binding of INTEGRATION.md section 2, as running code: subclasses of the REFERENCE's own
``ratinabox.Agent.Agent`` and ``ratinabox.Neurons.PlaceCells`` whose two overridable methods go to libriab_b200 through
its C ABI (``include/riab_b200.h``), nothing else changed -- construction, ``default_params``, the history lists and every
other method are the reference's.  One agent per object like the reference; the batched mirror classes of
``ratinabox_b200`` are the complete version of the same idea.

  Agent.update(dt, drift_velocity, drift_to_random_strength_ratio, **kwargs)   (ratinabox/Agent.py:160)   -> riab_agent_update
  PlaceCells.get_state(evaluate_at, **kwargs)                                  (ratinabox/Neurons.py:936) -> riab_place_rates

Needs the reference importable (``/root/reference`` or the copy staged by ``oracle/make_ref.py``; matplotlib / shapely may be
the stand-ins of ``oracle/ref_shim.py``) and a CUDA device.  ``tests/test_gpu_reference_binding.py`` runs it against the
unmodified classes."""
import ctypes as C

import numpy as np
import torch

from ratinabox.Agent import Agent
from ratinabox.Neurons import PlaceCells
from ratinabox_b200 import _lib                      # ctypes.Structure mirrors of riab_b200.h + prototypes

lib = _lib.load()                                    # CDLL("libriab_b200.so"); raises if it cannot load
DEV = torch.device("cuda")
_STATE = ("pos", "velocity", "rotational_velocity", "measured_velocity", "measured_rotational_velocity",
          "head_direction", "distance_travelled", "distance_to_closest_wall")
_stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def env_struct(Env, keep):
    walls = torch.as_tensor(np.ascontiguousarray(Env.walls, dtype=np.float64).reshape(-1, 4), device=DEV)
    keep.append(walls)
    e = _lib.Env()
    e.walls_dev, e.n_walls, e.n_boundary_walls = walls.data_ptr(), walls.shape[0], 4
    for i in range(4):
        e.extent[i] = float(Env.extent[i])
    e.boundary_mode, e.scale = (1 if Env.boundary_conditions == "periodic" else 0), float(getattr(Env, "scale", 1.0))
    return e


class CudaAgent(Agent):
    """The reference's Agent with update() on the GPU.  ``xi`` (two standard normals) may be injected for parity runs;
    otherwise the engine's Philox stream keyed on (seed, step) draws them."""

    def __init__(self, Environment, params={}, seed=0):
        super().__init__(Environment, params)
        self._seed, self._step, self._keep = seed, 0, []

    def update(self, dt=None, drift_velocity=None, drift_to_random_strength_ratio=1, xi=None, **kw):
        dt = dt or self.dt                                                        # Agent.py:193-194 (dt persists)
        self.dt = dt
        self.t += dt
        self.prev_pos, self.prev_velocity = self.pos.copy(), self.velocity.copy()
        self.prev_measured_velocity = self.measured_velocity.copy()
        # state -> device (one agent: a few doubles; the batched engine keeps it resident)
        host = {k: np.atleast_1d(np.asarray(getattr(self, k, 0.0 if k != "distance_to_closest_wall" else np.inf), dtype=np.float64))
                for k in _STATE}
        dev = {k: torch.as_tensor(np.ascontiguousarray(v.reshape(1, -1) if v.size == 2 else v.reshape(1)), device=DEV)
               for k, v in host.items()}
        ag = _lib.Agents()
        ag.n_agents, ag.id_offset = 1, 0
        for k in _STATE:
            setattr(ag, k, dev[k].data_ptr())
        mp = _lib.MotionParams()
        mp.dt = float(dt)
        mp.speed_mean, mp.speed_std, mp.speed_coherence_time = float(self.speed_mean), float(self.speed_std), float(self.speed_coherence_time)
        mp.speed_mean_kw = float(kw.get("speed_mean", self.speed_mean))
        mp.speed_coherence_time_kw = float(kw.get("speed_coherence_time", self.speed_coherence_time))
        mp.rotational_velocity_coherence_time_kw = float(kw.get("rotational_velocity_coherence_time", self.rotational_velocity_coherence_time))
        mp.rotational_velocity_std_kw = float(kw.get("rotational_velocity_std", self.rotational_velocity_std))
        mp.rotational_velocity_drift_kw = float(kw.get("rotational_velocity_drift", 0))
        mp.head_direction_smoothing_timescale = float(self.head_direction_smoothing_timescale)
        mp.thigmotaxis_kw = float(kw.get("thigmotaxis", self.thigmotaxis))
        mp.wall_repel_distance_kw = float(kw.get("wall_repel_distance", self.wall_repel_distance))
        mp.wall_repel_strength_kw = float(kw.get("wall_repel_strength", self.wall_repel_strength))
        mp.drift_to_random_strength_ratio = float(drift_to_random_strength_ratio)
        io = _lib.StepIO()
        io.seed, io.step = self._seed, self._step
        if drift_velocity is not None:
            drift = torch.as_tensor(np.asarray(drift_velocity, dtype=np.float64).reshape(1, 2), device=DEV)
            io.drift_velocity = drift.data_ptr()
        if xi is not None:
            tape = torch.as_tensor(np.asarray(xi, dtype=np.float64).reshape(1, 2), device=DEV)
            io.xi = tape.data_ptr()
        keep = []
        _lib.check(lib.riab_agent_update(C.byref(ag), C.byref(env_struct(self.Environment, keep)), C.byref(mp), C.byref(io), _stream()))
        self._step += 1
        for k in _STATE:                                                          # device -> the reference's attributes
            v = dev[k].cpu().numpy()
            setattr(self, k, v.reshape(2).copy() if v.size == 2 else float(v.reshape(-1)[0]))
        if self.save_history:
            self.save_to_history()                                                # the reference's own history lists


class CudaPlaceCells(PlaceCells):
    """The reference's PlaceCells with get_state() on the GPU (every description / wall geometry the engine has)."""

    def _cells(self, keep):
        env = self.Agent.Environment
        geom = _lib.WALL_GEOMETRIES[self.wall_geometry]
        centres = np.ascontiguousarray(self.place_cell_centres, dtype=np.float64)
        widths = np.ascontiguousarray(self.place_cell_widths, dtype=np.float64)
        walls = np.ascontiguousarray(env.walls, dtype=np.float64).reshape(-1, 4)
        extent = np.ascontiguousarray(env.extent, dtype=np.float64)
        n_inner = 0 if geom == 0 else len(walls) - 4
        meta = _lib.PlaceCells()
        meta.n_cells, meta.description, meta.wall_geometry = self.n, _lib.PC_DESCRIPTIONS[self.description], geom
        meta.min_fr, meta.max_fr, meta.top_hat_width = float(self.min_fr), float(self.max_fr), float(np.ravel(self.widths)[0])
        packed = np.zeros(lib.riab_place_pack_floats(self.n, n_inner), dtype=np.float32)
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        _lib.check(lib.riab_place_pack(centres.ctypes.data_as(dp), widths.ctypes.data_as(dp), self.n, walls.ctypes.data_as(dp),
                                       len(walls), 4, extent.ctypes.data_as(dp), geom, C.byref(meta), packed.ctypes.data_as(fp)))
        pk, ct = torch.as_tensor(packed, device=DEV), torch.as_tensor(centres, device=DEV)
        keep += [pk, ct]
        meta.packed_dev, meta.centres_dev = pk.data_ptr(), ct.data_ptr()
        return meta

    def get_state(self, evaluate_at="agent", **kwargs):
        if evaluate_at == "agent":
            pos = self.Agent.pos
        elif evaluate_at == "all":
            pos = self.Agent.Environment.flattened_discrete_coords
        else:
            pos = kwargs["pos"]
        pos = torch.as_tensor(np.ascontiguousarray(np.asarray(pos, dtype=np.float64).reshape(-1, 2)), device=DEV)
        keep = []
        ld = (self.n + 3) // 4 * 4
        out = torch.empty((pos.shape[0], ld), dtype=torch.float32, device=DEV)
        _lib.check(lib.riab_place_rates(pos.data_ptr(), pos.shape[0], C.byref(env_struct(self.Agent.Environment, keep)),
                                        C.byref(self._cells(keep)), out.data_ptr(), ld, _stream()))
        return out[:, : self.n].T.cpu().numpy().astype(np.float64)                # the reference's (n_cells, n_pos)
