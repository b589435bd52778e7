"""ctypes binding of libriab_b200.so (C ABI in include/riab_b200.h).

There is no CPU fallback: if the library cannot be loaded (or built in-tree with
nvcc) importing this module raises, and every entry point raises on a non-zero
status with the library's own message."""
import ctypes as C
import os

from . import _build

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


class RiabError(RuntimeError):
    pass


class Env(C.Structure):
    _fields_ = [("walls_dev", C.c_void_p), ("n_walls", C.c_int32), ("n_boundary_walls", C.c_int32),
                ("extent", C.c_double * 4), ("boundary_mode", C.c_int32), ("n_hole_walls", C.c_int32), ("scale", C.c_double),
                ("hole_wall0", C.c_int32), ("reserved", C.c_int32)]


class Agents(C.Structure):
    _fields_ = [("n_agents", C.c_int64), ("id_offset", C.c_int64), ("pos", C.c_void_p), ("velocity", C.c_void_p),
                ("rotational_velocity", C.c_void_p), ("measured_velocity", C.c_void_p),
                ("measured_rotational_velocity", C.c_void_p), ("head_direction", C.c_void_p),
                ("distance_travelled", C.c_void_p), ("distance_to_closest_wall", C.c_void_p)]


class MotionParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "dt", "speed_coherence_time_kw", "speed_mean_kw", "speed_mean", "speed_std", "speed_coherence_time",
        "rotational_velocity_coherence_time_kw", "rotational_velocity_std_kw", "rotational_velocity_drift_kw",
        "head_direction_smoothing_timescale", "thigmotaxis_kw", "wall_repel_distance_kw", "wall_repel_strength_kw",
        "drift_to_random_strength_ratio")]


class StepIO(C.Structure):
    _fields_ = [("drift_velocity", C.c_void_p), ("xi", C.c_void_p), ("seed", C.c_uint64), ("step", C.c_uint64),
                ("collision_mask", C.c_void_p), ("first_hit", C.c_void_p), ("n_iters", C.c_void_p),
                ("history_row", C.c_void_p), ("pos_mirror", C.c_void_p)]


class PlaceCells(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("description", C.c_int32), ("wall_geometry", C.c_int32),
                ("n_inner_walls", C.c_int32), ("min_fr", C.c_float), ("max_fr", C.c_float),
                ("top_hat_width", C.c_double), ("packed_dev", C.c_void_p), ("centres_dev", C.c_void_p),
                ("eps", C.c_float * 8), ("ep_valid", C.c_int32), ("n_pad", C.c_int32),
                ("k_uniform", C.c_float), ("r2_max", C.c_float)]


class GridCells(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("description", C.c_int32), ("width_ratio", C.c_double),
                ("min_fr", C.c_float), ("max_fr", C.c_float), ("packed_dev", C.c_void_p), ("n_pad", C.c_int32)]


class BvcCells(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_test_angles", C.c_int32), ("min_fr", C.c_float), ("max_fr", C.c_float),
                ("packed_dev", C.c_void_p), ("test_dirs_dev", C.c_void_p), ("n_pad", C.c_int32),
                ("egocentric", C.c_int32)]


MAX_OBJECTS = 9


class OvcCells(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_objects", C.c_int32), ("objects", C.c_double * (2 * MAX_OBJECTS)),
                ("object_types", C.c_int32 * MAX_OBJECTS), ("walls_occlude", C.c_int32), ("egocentric", C.c_int32),
                ("min_fr", C.c_float), ("max_fr", C.c_float), ("packed_dev", C.c_void_p), ("n_pad", C.c_int32),
                ("reserved", C.c_int32)]


class HistoryView(C.Structure):
    _fields_ = [("agent_ring", C.c_void_p), ("agent_ring_rows", C.c_int32), ("agent_row0", C.c_int32),
                ("rates_ring", C.c_void_p), ("rates_ring_rows", C.c_int32), ("rates_row0", C.c_int32),
                ("n_steps", C.c_int64), ("n_agents", C.c_int64), ("ld", C.c_int64), ("n_cells", C.c_int32),
                ("reserved", C.c_int32)]


class NeuronNoise(C.Structure):
    _fields_ = [("noise_std", C.c_float), ("noise_coherence_time", C.c_float), ("dt", C.c_float),
                ("seed", C.c_uint64), ("step", C.c_uint64), ("id_offset", C.c_int64), ("population_id", C.c_int32)]


class RatesOut(C.Structure):
    _fields_ = [("rates_row", C.c_void_p), ("ld", C.c_int64), ("spikes_row", C.c_void_p),
                ("noise_state", C.c_void_p), ("bvc_scratch", C.c_void_p)]


class Population(C.Structure):
    _fields_ = [("kind", C.c_int32), ("cells", C.c_void_p), ("noise", NeuronNoise), ("out", RatesOut),
                ("rates_ring", C.c_void_p), ("spikes_ring", C.c_void_p), ("ring_rows", C.c_int32),
                ("ring_next", C.c_int32)]


class AgentHistory(C.Structure):
    _fields_ = [("ring", C.c_void_p), ("ring_rows", C.c_int32), ("ring_next", C.c_int32)]


PC_DESCRIPTIONS = {"gaussian": 0, "gaussian_threshold": 1, "diff_of_gaussians": 2, "top_hat": 3, "one_hot": 4}
WALL_GEOMETRIES = {"euclidean": 0, "line_of_sight": 1, "geodesic": 2}
GC_DESCRIPTIONS = {"rectified_cosines": 0, "shifted_cosines": 1}
CELLS_PLACE, CELLS_GRID, CELLS_BVC, CELLS_OVC = 0, 1, 2, 3
MAX_REC_ITERS = 4

# name -> (restype, argtypes); every symbol include/riab_b200.h declares
SYMBOLS = {
    "riab_abi_version": (C.c_int, []),
    "riab_last_error": (C.c_char_p, []),
    "riab_launch_count": (C.c_int64, []),
    "riab_stream_synchronize": (C.c_int, [C.c_void_p]),
    "riab_history_rate_maps": (C.c_int, [C.POINTER(HistoryView), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "riab_agent_update": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.POINTER(MotionParams), C.POINTER(StepIO), C.c_void_p]),
    "riab_place_pack_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "riab_place_pack": (C.c_int, [c_double_p, c_double_p, C.c_int32, c_double_p, C.c_int32, C.c_int32, c_double_p,
                                  C.c_int32, C.POINTER(PlaceCells), c_float_p]),
    "riab_place_rates": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Env), C.POINTER(PlaceCells), C.c_void_p, C.c_int64, C.c_void_p]),
    "riab_grid_pack_floats": (C.c_int64, [C.c_int32]),
    "riab_grid_pack": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int32, c_double_p, C.POINTER(GridCells), c_float_p]),
    "riab_grid_rates": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Env), C.POINTER(GridCells), C.c_void_p, C.c_int64, C.c_void_p]),
    "riab_bvc_pack_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "riab_bvc_scratch_floats": (C.c_int64, [C.c_int64, C.c_int32]),
    "riab_bvc_pack": (C.c_int, [c_double_p, c_double_p, c_double_p, c_double_p, C.c_int32, c_double_p, C.c_int32,
                                C.POINTER(BvcCells), c_float_p]),
    "riab_bvc_rates": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Env), C.POINTER(BvcCells), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "riab_ovc_pack_floats": (C.c_int64, [C.c_int32]),
    "riab_ovc_pack": (C.c_int, [c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int32), C.c_int32,
                                C.POINTER(OvcCells), c_float_p]),
    "riab_ovc_rates": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Env), C.POINTER(OvcCells), C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_void_p]),
    "riab_step_fused": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.POINTER(MotionParams), C.POINTER(StepIO),
                                  C.c_int32, C.c_void_p, C.POINTER(NeuronNoise), C.POINTER(RatesOut), C.c_void_p]),
    "riab_neurons_update": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.c_int32, C.c_void_p, C.POINTER(NeuronNoise),
                                      C.POINTER(RatesOut), C.c_void_p]),
    "riab_run": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.POINTER(MotionParams), C.POINTER(StepIO),
                           C.POINTER(Population), C.c_int32, C.POINTER(AgentHistory), C.c_int64, C.c_void_p]),
    "riab_agent_update_host": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.POINTER(MotionParams), C.POINTER(StepIO),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "riab_positions_wait": (C.c_int, []),
    "riab_positions_fence": (C.c_int, [C.c_void_p]),
    "riab_step_fused_host": (C.c_int, [C.POINTER(Agents), C.POINTER(Env), C.POINTER(MotionParams), C.POINTER(StepIO),
                                       C.c_int32, C.c_void_p, C.POINTER(NeuronNoise), C.POINTER(RatesOut),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building in-tree first if needed).  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not _have_nvcc():
            raise ImportError(f"{path} is missing and nvcc is not available to build it: "
                              "ratinabox_b200 has no CPU fallback")
        _build.build()
    elif os.path.isdir(_build.CSRC) and _build.needs_build():
        # sources newer than the library (or file times scrambled by a copy): rebuilding takes minutes, so it is only
        # done on request (__graft_entry__.build(), `python ratinabox_b200/_build.py`, RIAB_AUTO_REBUILD=1)
        if os.environ.get("RIAB_AUTO_REBUILD") == "1" and _have_nvcc():
            _build.build()
        else:
            import warnings
            warnings.warn(f"{path} is older than its CUDA sources; run __graft_entry__.build() to rebuild")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.riab_abi_version() != 2:
        raise ImportError("libriab_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _have_nvcc():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return os.path.exists(nvcc)


def check(rc):
    if rc != 0:
        raise RiabError(f"libriab_b200 status {rc}: {load().riab_last_error().decode()}")
