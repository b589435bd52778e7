"""Pin the CPU oracle (oracle/riab_oracle.py) against fixtures produced by the
LIVE unmodified reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import riab_oracle as O


def _restore_rng(g):
    np.random.set_state(("MT19937", g["rng_keys"], int(g["rng_pos"]), int(g["rng_has_gauss"]),
                         float(g["rng_cached"])))


def test_native_c1_bit_exact(golden):
    """Config 1, global RNG, jitter ON: the oracle consumes the RNG in the
    reference's order and must reproduce its history bit for bit."""
    g = golden("native_c1.npz")
    env = O.OracleEnvironment()
    ag = O.OracleAgent(env, g["pos0"], g["vel0"], {"dt": 0.01})
    centres, widths = g["centres"], g["widths"]
    rng = O.GlobalRNG()
    pcs = O.OracleNeurons(ag, len(widths), lambda pos, r: O.place_cells_get_state(
        env, centres, widths, pos, r, "gaussian", "euclidean"))
    assert str(g["wall_geometry"]) == "geodesic"      # W == 4 -> plain euclidean distances
    _restore_rng(g)
    for _ in range(600):
        ag.update(rng)
        pcs.update(rng)
    for key, hk in (("pos", "pos"), ("vel", "vel"), ("rot_vel", "rot_vel"), ("head_direction", "head_direction"),
                    ("distance_travelled", "distance_travelled"), ("t", "t")):
        assert np.array_equal(np.array(ag.history[hk]), g[key]), key
    assert np.array_equal(np.array(pcs.history["firingrate"]), g["firingrate"])
    spikes = np.unpackbits(g["spikes"])[: 600 * 100].reshape(600, 100).astype(bool)
    assert np.array_equal(np.array(pcs.history["spikes"]), spikes)


def test_native_walls_bit_exact(golden):
    """2x1 box, two internal walls, fast agent that bounces; line_of_sight
    gaussian_threshold PlaceCells + GridCells + BVCs; global RNG, jitter ON."""
    g = golden("native_walls.npz")
    env = O.OracleEnvironment(scale=1, aspect=2, walls=[[[1, 0], [1, 0.35]], [[1, 0.65], [1, 1]]])
    assert np.array_equal(env.walls, g["walls"])
    ag = O.OracleAgent(env, g["pos0"], g["vel0"], {"dt": 0.05, "speed_mean": 0.4})
    rng = O.GlobalRNG()
    pcs = O.OracleNeurons(ag, 20, lambda pos, r: O.place_cells_get_state(
        env, g["centres"], g["widths"], pos, r, "gaussian_threshold", "line_of_sight", 0.1, 10))
    gcs = O.OracleNeurons(ag, 12, lambda pos, r: O.grid_cells_get_state(
        g["gridscales"], g["phase_offsets"], g["gc_w"], pos))
    bvcs = O.OracleNeurons(ag, 10, lambda pos, r: O.bvc_get_state(
        env, g["bvc_mu_d"], g["bvc_mu_t"], g["bvc_sg_d"], g["bvc_sg_t"], pos, r))
    dirs, angs = O.bvc_test_angles(2)
    assert np.array_equal(dirs, g["bvc_test_dirs"]) and np.array_equal(angs, g["bvc_test_angles"])
    assert np.array_equal(O.bvc_cell_fr_norm(angs, g["bvc_sg_t"]), g["bvc_norm"])
    _restore_rng(g)
    nb = 0
    for _ in range(1500):
        info = ag.update(rng)
        nb += len(info["first_hit"]) > 0
        pcs.update(rng)
        gcs.update(rng)
        bvcs.update(rng)
    assert nb == int(g["n_bounces"]) and nb > 0
    assert np.array_equal(np.array(ag.history["pos"]), g["pos"])
    assert np.array_equal(np.array(ag.history["vel"]), g["vel"])
    assert np.array_equal(np.array(ag.history["rot_vel"]), g["rot_vel"])
    assert np.array_equal(np.array(ag.history["head_direction"]), g["head_direction"])
    assert np.array_equal(np.array(ag.history["distance_travelled"]), g["distance_travelled"])
    assert np.array_equal(np.array(pcs.history["firingrate"]), g["pc_fr"])
    assert np.array_equal(np.array(gcs.history["firingrate"]), g["gc_fr"])
    assert np.array_equal(np.array(bvcs.history["firingrate"]), g["bvc_fr"])


@pytest.mark.parametrize("name", ["box2", "maze8"])
def test_modeA_motion(golden, name):
    """Teacher-forced single steps, zero jitter, injected normals."""
    g = golden(f"modeA_motion_{name}.npz")
    env = O.OracleEnvironment(walls=g["walls"][4:])
    assert np.array_equal(env.walls, g["walls"])
    A = len(g["pos0"])
    shp = tuple(g["masks_shape"])
    masks = np.unpackbits(g["masks"])[: int(np.prod(shp))].reshape(shp).astype(bool)
    for a in range(A):
        ag = O.OracleAgent(env, g["pos0"][a], g["vel0"][a], {"dt": 0.01})
        ag.rotational_velocity = float(g["rot0"][a])
        ag.measured_velocity = g["mv0"][a].copy()
        ag.head_direction = g["hd0"][a].copy()
        ag.distance_travelled = float(g["dist0"][a])
        rng = O.TapeRNG(agent_xi=g["xi"][a])
        if g["use_drift"][a]:
            info = ag.update(rng, drift_velocity=g["drift"][a], drift_to_random_strength_ratio=float(g["drift_ratio"]))
        else:
            info = ag.update(rng)
        assert np.array_equal(ag.pos, g["out_pos"][a]), a
        assert np.array_equal(ag.velocity, g["out_vel"][a]), a
        assert ag.rotational_velocity == g["out_rot"][a]
        assert np.array_equal(ag.measured_velocity, g["out_mv"][a])
        assert ag.measured_rotational_velocity == g["out_mrot"][a]
        assert np.array_equal(ag.head_direction, g["out_hd"][a])
        assert ag.distance_travelled == g["out_dist"][a]
        assert ag.distance_to_closest_wall == g["out_dclose"][a]
        assert len(info["collisions"]) == g["out_n_iter"][a]
        for i, m in enumerate(info["collisions"][:4]):
            assert np.array_equal(m, masks[a, i])
        assert (info["first_hit"] + [-1] * 4)[:4] == list(g["out_first_hit"][a])


def test_modeA_rates(golden):
    g = golden("modeA_rates.npz")
    P = g["P"]
    rng = O.TapeRNG()
    env = O.OracleEnvironment(walls=g["box2_walls"][4:])
    for desc in ("gaussian", "gaussian_threshold", "diff_of_gaussians", "top_hat", "one_hot"):
        for geom in ("euclidean", "line_of_sight"):
            c = g[f"pc_{desc}_{geom}_centres"]
            fr = O.place_cells_get_state(env, c, 0.2 * np.ones(len(c)), P, rng, desc, geom, 0.05, 3.0, scalar_width=0.2)
            assert np.array_equal(fr, g[f"pc_{desc}_{geom}"]), (desc, geom)
    env1 = O.OracleEnvironment(walls=g["geodesic_walls"][4:])
    c = g["pc_gaussian_geodesic_centres"]
    fr = O.place_cells_get_state(env1, c, 0.15 * np.ones(len(c)), P, rng, "gaussian", "geodesic")
    assert np.array_equal(fr, g["pc_gaussian_geodesic"])
    for desc in ("rectified_cosines", "shifted_cosines"):
        w = O.grid_cells_w(g[f"gc_{desc}_orient"])
        assert np.array_equal(w, g[f"gc_{desc}_w"])
        fr = O.grid_cells_get_state(g[f"gc_{desc}_gridscales"], g[f"gc_{desc}_phase"], w, P, desc,
                                    min_fr=0.1, max_fr=2.0)
        assert np.array_equal(fr, g[f"gc_{desc}"]), desc
    for name in ("box2", "maze8"):
        e = O.OracleEnvironment(walls=g[f"bvc_{name}_walls"][4:])
        fr = O.bvc_get_state(e, g[f"bvc_{name}_tuning_distances"], g[f"bvc_{name}_tuning_angles"],
                             g[f"bvc_{name}_sigma_distances"], g[f"bvc_{name}_sigma_angles"], P, rng,
                             min_fr=0.0, max_fr=5.0)
        assert np.array_equal(fr, g[f"bvc_{name}"]), name


def test_modeA_fov_egocentric_bvcs(golden):
    """FieldOfViewBVCs (egocentric frame, diverging manifold) -- Neurons.py:1693-1708, :1847-1887,
    utils.py:1073-1112 -- against the live reference."""
    g = golden("modeA_fov.npz")
    P, HD = g["P"], g["HD"]
    mu_d, mu_t, sg_d, sg_t = O.diverging_radial_assembly(distance_range=[0.02, 0.4], angle_range=[0, 75],
                                                         spatial_resolution=0.02, beta=5)
    rng = O.TapeRNG()
    for name in ("box2", "maze8"):
        assert np.array_equal(mu_d, g[f"fov_{name}_tuning_distances"])
        assert np.array_equal(mu_t, g[f"fov_{name}_tuning_angles"])
        assert np.array_equal(sg_d, g[f"fov_{name}_sigma_distances"])
        assert np.array_equal(sg_t, g[f"fov_{name}_sigma_angles"])
        env = O.OracleEnvironment(walls=g[f"fov_{name}_walls"][4:])
        ref = g[f"fov_{name}"]
        for j in range(0, len(P), 3):
            fr = O.bvc_get_state(env, mu_d, mu_t, sg_d, sg_t, P[j], rng, min_fr=0.0, max_fr=2.0, head_direction=HD[j])
            assert np.array_equal(fr[:, 0], ref[:, j]), (name, j)
        hd0 = HD[0] / np.linalg.norm(HD[0])
        fr = O.bvc_get_state(env, mu_d, mu_t, sg_d, sg_t, P[0], rng, min_fr=0.0, max_fr=2.0, head_direction=hd0)
        assert np.array_equal(fr[:, 0], g[f"fov_{name}_agent"])


@pytest.mark.parametrize("name,walls", [("open", []), ("wall", [[[0.5, 0.2], [0.5, 0.8]]])])
def test_periodic_boundary_conditions(golden, name, walls):
    """Periodic rectangular box (Environment.py:130-136, :670-675, :877-879): native 800-step run (global RNG,
    jitter on, 26-35 boundary crossings) bit for bit, and teacher-forced single steps across the boundary."""
    g = golden("periodic.npz")
    env = O.OracleEnvironment(walls=walls, boundary_conditions="periodic")
    assert np.array_equal(env.walls, g[f"{name}_walls"].reshape(-1, 2, 2))
    assert str(g[f"{name}_geom"]) == "euclidean"
    ag = O.OracleAgent(env, g[f"{name}_pos0"], g[f"{name}_vel0"], {"dt": 0.05, "speed_mean": 0.5})
    rng = O.GlobalRNG()
    pcs = O.OracleNeurons(ag, 30, lambda p, r: O.place_cells_get_state(env, g[f"{name}_centres"], g[f"{name}_widths"], p, r))
    gcs = O.OracleNeurons(ag, 9, lambda p, r: O.grid_cells_get_state(g[f"{name}_gridscales"], g[f"{name}_phase"], g[f"{name}_w"], p))
    np.random.set_state(("MT19937", g[f"{name}_rng_keys"], int(g[f"{name}_rng_pos"]), int(g[f"{name}_rng_has_gauss"]),
                         float(g[f"{name}_rng_cached"])))
    for _ in range(800):
        ag.update(rng); pcs.update(rng); gcs.update(rng)
    assert np.array_equal(np.array(ag.history["pos"]), g[f"{name}_pos"])
    assert np.array_equal(np.array(ag.history["vel"]), g[f"{name}_vel"])
    assert np.array_equal(np.array(ag.history["rot_vel"]), g[f"{name}_rot_vel"])
    assert np.array_equal(np.array(ag.history["distance_travelled"]), g[f"{name}_dist"])
    assert np.array_equal(np.array(pcs.history["firingrate"]), g[f"{name}_pc_fr"])
    assert np.array_equal(np.array(gcs.history["firingrate"]), g[f"{name}_gc_fr"])
    for a in range(len(g[f"{name}_A_pos0"])):
        v0 = g[f"{name}_A_vel0"][a]
        oa = O.OracleAgent(env, g[f"{name}_A_pos0"][a], v0, {"dt": 0.05, "speed_mean": 0.5})
        oa.update(O.TapeRNG(agent_xi=g[f"{name}_A_xi"][a]))
        assert np.array_equal(oa.pos, g[f"{name}_A_pos"][a]) and np.array_equal(oa.measured_velocity, g[f"{name}_A_mv"][a])
        assert oa.distance_travelled == g[f"{name}_A_dist"][a]
    fr = O.place_cells_get_state(env, g[f"{name}_centres"], g[f"{name}_widths"], g[f"{name}_A_pos0"], O.TapeRNG())
    assert np.array_equal(fr, g[f"{name}_A_pc"])


POLY_CASES = {"lroom": dict(boundary=[[0, 0], [1, 0], [1, 0.5], [0.5, 0.5], [0.5, 1], [0, 1]], walls=[[[0.25, 0.0], [0.25, 0.3]]]),
              "holed": dict(holes=[[[0.4, 0.4], [0.6, 0.4], [0.6, 0.6], [0.4, 0.6]]], walls=[[[0.8, 0.0], [0.8, 0.35]]])}


@pytest.mark.parametrize("name", sorted(POLY_CASES))
def test_polygon_boundary_and_holes(golden, name):
    """Polygon boundary / holes (Environment.py:112-160, :807-817): wall list, a native 1000-step run (global RNG,
    jitter on) with line_of_sight PlaceCells (the hard-coded `walls[4:]`, Environment.py:715-717) and BVCs bit for
    bit, and 384 teacher-forced single steps started next to boundary / hole / inner walls."""
    g = golden("polygon.npz")
    env = O.OracleEnvironment(**POLY_CASES[name])
    assert np.array_equal(env.walls, g[f"{name}_walls"]) and np.array_equal(env.extent, g[f"{name}_extent"])
    assert str(g[f"{name}_geom"]) == "line_of_sight"
    prm = {"dt": 0.02, "speed_mean": 0.25}
    ag = O.OracleAgent(env, g[f"{name}_pos0"], g[f"{name}_vel0"], prm)
    rng = O.GlobalRNG()
    td, ta, sd, sa = g[f"{name}_bvc"]
    pcs = O.OracleNeurons(ag, 24, lambda p, r: O.place_cells_get_state(env, g[f"{name}_centres"], g[f"{name}_widths"], p, r,
                                                                       "gaussian", "line_of_sight"))
    bvcs = O.OracleNeurons(ag, 6, lambda p, r: O.bvc_get_state(env, td, ta, sd, sa, p, r))
    np.random.set_state(("MT19937", g[f"{name}_rng_keys"], int(g[f"{name}_rng_pos"]), int(g[f"{name}_rng_has_gauss"]),
                         float(g[f"{name}_rng_cached"])))
    for _ in range(1000):
        ag.update(rng); pcs.update(rng); bvcs.update(rng)
    assert np.array_equal(np.array(ag.history["pos"]), g[f"{name}_pos"])
    assert np.array_equal(np.array(ag.history["vel"]), g[f"{name}_vel"])
    assert np.array_equal(np.array(pcs.history["firingrate"]), g[f"{name}_pc_fr"])
    assert np.array_equal(np.array(bvcs.history["firingrate"]), g[f"{name}_bvc_fr"])
    assert all(env.contains(p) for p in g[f"{name}_pos"][::10])
    for a in range(len(g[f"{name}_A_pos0"])):
        assert env.contains(g[f"{name}_A_pos0"][a])
        oa = O.OracleAgent(env, g[f"{name}_A_pos0"][a], g[f"{name}_A_vel0"][a], prm)
        oa.update(O.TapeRNG(agent_xi=g[f"{name}_A_xi"][a]))
        assert np.array_equal(oa.pos, g[f"{name}_A_pos"][a]) and np.array_equal(oa.velocity, g[f"{name}_A_vel"][a])
        assert np.array_equal(oa.measured_velocity, g[f"{name}_A_mv"][a])
    fr = O.place_cells_get_state(env, g[f"{name}_centres"], g[f"{name}_widths"], g[f"{name}_A_pos0"], O.TapeRNG(),
                                 "gaussian", "line_of_sight")
    assert np.array_equal(fr, g[f"{name}_A_pc"])
    assert np.array_equal(O.bvc_get_state(env, td, ta, sd, sa, g[f"{name}_A_pos0"], O.TapeRNG()), g[f"{name}_A_bvc"])


OVC_WALLS = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]


def test_object_vector_cells(golden):
    """ObjectVectorCells / FieldOfViewOVCs (Neurons.py:1892-2160) against the live reference: a native 400-step
    run (global RNG, jitter on; allocentric + euclidean + egocentric field-of-view populations on one Agent) bit
    for bit, and get_state at 384 positions / head directions in zero-jitter mode."""
    g = golden("ovc.npz")
    env = O.OracleEnvironment(walls=OVC_WALLS)
    obj, otypes = g["objects"], g["object_types"]
    assert list(otypes) == [0, 0, 1, 2, 1]
    ag = O.OracleAgent(env, g["pos0"], g["vel0"], {"dt": 0.02})
    rng = O.GlobalRNG()

    def pop(k, ego):
        td, ta, sd, sa = g[f"{k}_tuning"]
        geom = str(g[f"{k}_geom"])
        return O.OracleNeurons(ag, len(td), lambda p, r: O.ovc_get_state(
            env, obj, otypes, td, ta, sd, sa, g[f"{k}_types"], p, r, geom, head_direction=(ag.head_direction if ego else None)))

    pops = {"allo": pop("allo", False), "eucl": pop("eucl", False), "fov": pop("fov", True)}
    np.random.set_state(("MT19937", g["rng_keys"], int(g["rng_pos"]), int(g["rng_has_gauss"]), float(g["rng_cached"])))
    for _ in range(400):
        ag.update(rng)
        for P in pops.values():
            P.update(rng)
    assert np.array_equal(np.array(ag.history["pos"]), g["pos"])
    for k, P in pops.items():
        assert np.array_equal(np.array(P.history["firingrate"]), g[f"{k}_fr"]), k
    for k, ego in (("allo", False), ("eucl", False), ("fov", True)):
        td, ta, sd, sa = g[f"{k}_tuning"]
        fr = O.ovc_get_state(env, obj, otypes, td, ta, sd, sa, g[f"{k}_types"], g["A_pos"], O.TapeRNG(), str(g[f"{k}_geom"]),
                             head_direction=(g["A_hd"] if ego else None))
        assert np.array_equal(fr, g[f"A_{k}"]), k
    assert (g["A_allo"] > 0.05).mean() > 0.01 and (g["A_fov"] > 0.05).mean() > 0.002


BOX_WALLS = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]

PARAM_VARIANTS = {          # (constructor params, update kwargs); same table as oracle/gen_golden.py
    "speed_std0": ({"speed_std": 0.0}, {}),
    "thigmotaxis0": ({"thigmotaxis": 0.0}, {}),
    "thigmotaxis1": ({"thigmotaxis": 1.0}, {}),
    "no_repel": ({"wall_repel_strength": 0.0}, {}),
    "strong_repel": ({"wall_repel_strength": 2.5, "wall_repel_distance": 0.2}, {}),
    "fast_head": ({"head_direction_smoothing_timescale": 0.005}, {}),
    "timescales": ({"speed_coherence_time": 0.1, "rotational_velocity_coherence_time": 0.3,
                    "rotational_velocity_std": 3.0, "speed_mean": 0.2}, {}),
    "kwargs": ({}, {"speed_mean": 0.3, "speed_coherence_time": 0.2, "rotational_velocity_std": 1.0,
                    "rotational_velocity_coherence_time": 0.05, "rotational_velocity_drift": 0.7,
                    "thigmotaxis": 0.8, "wall_repel_distance": 0.15, "wall_repel_strength": 1.5,
                    "head_direction_smoothing_timescale": 0.4}),
    "kw_std0": ({"speed_std": 0.0}, {"speed_mean": 0.3}),
    "dt_arg": ({}, {"dt": 0.05}),
    "drift_weak": ({}, {"drift_ratio": 0.5}),
    "drift_strong": ({"speed_mean": 0.15}, {"drift_ratio": 5.0}),
}


@pytest.mark.parametrize("name", sorted(PARAM_VARIANTS))
def test_agent_parameter_and_keyword_variants(golden, name):
    """Every motion parameter / Agent.update keyword the reference reads (Agent.py:280-285, :302-311, :353-355,
    :483-489), including its quirks -- the attribute speed_std with the kwarg speed_mean, the overwritten
    head_direction_smoothing_timescale kwarg, dt that persists: 160 teacher-forced steps per variant, bit for bit."""
    g = golden("modeA_params.npz")
    params, kw = PARAM_VARIANTS[name]
    kw = dict(kw)
    ratio = kw.pop("drift_ratio", None)
    env = O.OracleEnvironment(walls=BOX_WALLS)
    assert np.array_equal(env.walls, g["walls"])
    for a in range(len(g["pos0"])):
        oa = O.OracleAgent(env, g["pos0"][a], g["vel0"][a], dict({"dt": 0.01}, **params))
        oa.rotational_velocity, oa.measured_velocity = float(g["rot0"][a]), g["mv0"][a].copy()
        oa.head_direction = g["hd0"][a].copy()
        rng = O.TapeRNG(agent_xi=g["xi"][a])
        if ratio is not None:
            oa.update(rng, drift_velocity=g["drift"][a].copy(), drift_to_random_strength_ratio=ratio, **kw)
        else:
            oa.update(rng, **kw)
        for key, val in (("pos", oa.pos), ("vel", oa.velocity), ("rot", oa.rotational_velocity), ("mv", oa.measured_velocity),
                         ("mrot", oa.measured_rotational_velocity), ("hd", oa.head_direction), ("dist", oa.distance_travelled),
                         ("dclose", oa.distance_to_closest_wall)):
            assert np.array_equal(np.asarray(val), g[f"{name}_{key}"][a]), (name, a, key)
        if name == "dt_arg":
            assert oa.dt == 0.05


def test_bin_data_for_histogramming(golden):
    """utils.bin_data_for_histogramming (utils.py:544-589) -- the oracle's restatement and the host mirror in
    ratinabox_b200.utils against the live reference: plain, weighted, bin-count-normalised, zero-bin mask; samples on
    bin edges, on the right-most edges and outside the extent."""
    from ratinabox_b200 import utils as U
    g = golden("histogram.npz")
    data, w, extent, dx = g["data"], g["weights"], list(g["extent"]), float(g["dx"])
    for f in (O.bin_data_for_histogramming, U.bin_data_for_histogramming):
        assert np.array_equal(f(data, extent, dx), g["plain"])
        assert np.array_equal(f(data, extent, dx, weights=w), g["weighted"])
        hm, zb = f(data, extent, dx, weights=w, norm_by_bincount=True, return_zero_bins=True)
        assert np.array_equal(hm, g["normed"]) and np.array_equal(zb, g["zero_bins"])
