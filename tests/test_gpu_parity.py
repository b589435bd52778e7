"""Parity of the CUDA path (through the ratinabox_b200 Python mirror -> C ABI) against
the golden fixtures of the live reference and against the CPU oracle.  GPU only.

Tolerances (BASELINE.json north_star): positions <= 1e-6 m, firing rates <= 1e-5
relative (evaluated relative to the population's rate scale max_fr-min_fr, plus a
pure relative check on entries above 1e-3 of that scale), wall-collision masks and
first-hit indices bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_TOL = 1e-6
RATE_TOL = 1e-5


def _env(rb, walls):
    E = rb.Environment()
    for w in walls[4:]:
        E.add_wall(w)
    assert np.array_equal(E.walls, walls)
    return E


def assert_rates_close(got, ref, scale, what="", pure_rel=True):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    assert err.max() <= RATE_TOL * scale, f"{what}: max abs err {err.max():.3e} > {RATE_TOL * scale:.1e}"
    big = np.abs(ref) > 1e-3 * scale
    if pure_rel and big.any():
        rel = (err[big] / np.abs(ref[big])).max()
        assert rel <= RATE_TOL, f"{what}: max rel err {rel:.3e}"


@pytest.mark.parametrize("name", ["box2", "maze8"])
def test_motion_modeA_golden(golden, name):
    import ratinabox_b200 as rb
    g = golden(f"modeA_motion_{name}.npz")
    shp = tuple(g["masks_shape"])
    masks = np.unpackbits(g["masks"])[: int(np.prod(shp))].reshape(shp).astype(bool)
    for with_drift in (False, True):
        sel = np.where(g["use_drift"] == with_drift)[0]
        E = _env(rb, g["walls"])
        Ag = rb.Agent(E, {"dt": 0.01, "n_agents": len(sel)})
        Ag.pos, Ag.velocity = g["pos0"][sel], g["vel0"][sel]
        Ag.rotational_velocity = g["rot0"][sel]
        Ag.measured_velocity, Ag.head_direction = g["mv0"][sel], g["hd0"][sel]
        Ag.distance_travelled = g["dist0"][sel]
        kw = dict(_xi=g["xi"][sel], _record_collisions=True)
        if with_drift:
            Ag.update(drift_velocity=g["drift"][sel], drift_to_random_strength_ratio=float(g["drift_ratio"]), **kw)
        else:
            Ag.update(**kw)
        info = Ag.last_collision_info()
        assert np.abs(Ag.pos - g["out_pos"][sel]).max() <= 1e-12          # far inside the 1e-6 m tolerance
        assert np.abs(Ag.velocity - g["out_vel"][sel]).max() <= 1e-12
        assert np.abs(Ag.rotational_velocity - g["out_rot"][sel]).max() <= 1e-10
        assert np.abs(Ag.measured_velocity - g["out_mv"][sel]).max() <= 1e-10
        assert np.abs(Ag.measured_rotational_velocity - g["out_mrot"][sel]).max() <= 1e-7
        assert np.abs(Ag.head_direction - g["out_hd"][sel]).max() <= 1e-12
        assert np.abs(Ag.distance_travelled - g["out_dist"][sel]).max() <= 1e-12
        assert np.abs(Ag.distance_to_closest_wall - g["out_dclose"][sel]).max() <= 1e-12
        # wall-collision indices: bit-exact
        assert np.array_equal(info["n_iters"], g["out_n_iter"][sel])
        assert np.array_equal(info["first_hit"], g["out_first_hit"][sel])
        assert np.array_equal(info["mask"].astype(bool), masks[sel])
        assert (g["out_n_iter"][sel] > 1).sum() > 0          # the case set does contain bounces
        h = Ag.get_history_arrays()
        assert np.abs(h["pos"][-1] - g["out_pos"][sel]).max() <= POS_TOL


def test_place_cells_modeA_golden(golden):
    import ratinabox_b200 as rb
    g = golden("modeA_rates.npz")
    P = g["P"]
    E = _env(rb, g["box2_walls"])
    Ag = rb.Agent(E, {"dt": 0.01})
    for desc in ("gaussian", "gaussian_threshold", "diff_of_gaussians", "top_hat", "one_hot"):
        for geom in ("euclidean", "line_of_sight"):
            c = g[f"pc_{desc}_{geom}_centres"]
            pc = rb.PlaceCells(Ag, {"place_cell_centres": c, "description": desc, "wall_geometry": geom,
                                    "widths": 0.2, "min_fr": 0.05, "max_fr": 3.0})
            got = pc.get_state(evaluate_at=None, pos=P)
            ref = g[f"pc_{desc}_{geom}"]
            if desc in ("top_hat", "one_hot"):
                # discontinuous profile: the in/out classification must be identical, values float32-rounded
                assert np.array_equal(got > 1.5, ref > 1.5), (desc, geom)
                assert np.abs(got - ref).max() <= 1e-6
            else:
                # diff_of_gaussians is a difference of two O(1) terms: near its zero crossing a pure
                # relative criterion is ill-posed in float32, so it is held to 1e-5 of the rate scale
                assert_rates_close(got, ref, 3.0 - 0.05, f"pc {desc} {geom}", pure_rel=(desc == "gaussian"))
    # geodesic: one internal wall
    E1 = _env(rb, g["geodesic_walls"])
    Ag1 = rb.Agent(E1, {"dt": 0.01})
    c = g["pc_gaussian_geodesic_centres"]
    pc = rb.PlaceCells(Ag1, {"place_cell_centres": c, "wall_geometry": "geodesic", "widths": 0.15})
    assert_rates_close(pc.get_state(evaluate_at=None, pos=P), g["pc_gaussian_geodesic"], 1.0, "pc geodesic")


def test_grid_cells_modeA_golden(golden):
    import ratinabox_b200 as rb
    g = golden("modeA_rates.npz")
    P = g["P"]
    E = _env(rb, g["box2_walls"])
    Ag = rb.Agent(E, {"dt": 0.01})
    for desc in ("rectified_cosines", "shifted_cosines"):
        gc = rb.GridCells(Ag, {"gridscale": g[f"gc_{desc}_gridscales"], "orientation": g[f"gc_{desc}_orient"],
                               "phase_offset": g[f"gc_{desc}_phase"], "description": desc,
                               "min_fr": 0.1, "max_fr": 2.0})
        assert np.array_equal(gc.w, g[f"gc_{desc}_w"])
        # sums / rectified differences of cosines: held to 1e-5 of the rate scale (no pure relative check)
        assert_rates_close(gc.get_state(evaluate_at=None, pos=P), g[f"gc_{desc}"], 1.9, f"gc {desc}", pure_rel=False)


@pytest.mark.parametrize("name", ["box2", "maze8"])
def test_bvc_modeA_golden(golden, name):
    import ratinabox_b200 as rb
    g = golden("modeA_rates.npz")
    P = g["P"]
    E = _env(rb, g[f"bvc_{name}_walls"])
    Ag = rb.Agent(E, {"dt": 0.01})
    bvc = rb.BoundaryVectorCells(Ag, {
        "tuning_distance": g[f"bvc_{name}_tuning_distances"], "tuning_angle": np.degrees(g[f"bvc_{name}_tuning_angles"]),
        "sigma_distance": g[f"bvc_{name}_sigma_distances"], "sigma_angle": np.degrees(g[f"bvc_{name}_sigma_angles"]),
        "min_fr": 0.0, "max_fr": 5.0})
    # degrees -> radians round trip is not exact: pin the exact reference values
    bvc.tuning_angles, bvc.sigma_angles = g[f"bvc_{name}_tuning_angles"], g[f"bvc_{name}_sigma_angles"]
    assert np.array_equal(bvc.test_angles, g[f"bvc_{name}_test_angles"])
    assert np.array_equal(bvc.test_directions, g[f"bvc_{name}_test_directions"])
    assert_rates_close(bvc.get_state(evaluate_at=None, pos=P), g[f"bvc_{name}"], 5.0, f"bvc {name}")


@pytest.mark.parametrize("name", ["box2", "maze8"])
def test_field_of_view_bvcs_egocentric_golden(golden, name):
    """FieldOfViewBVCs (egocentric frame; Neurons.py:1693-1708, :1847-1887) against the live reference:
    same manifold, rates at 160 positions each with its own head direction, and the Agent's own."""
    import ratinabox_b200 as rb
    g = golden("modeA_fov.npz")
    P, HD = g["P"], g["HD"]
    E = _env(rb, g[f"fov_{name}_walls"])
    Ag = rb.Agent(E, {"dt": 0.01})
    fov = rb.FieldOfViewBVCs(Ag, {"min_fr": 0.0, "max_fr": 2.0})
    for k, attr in (("tuning_distances", "tuning_distances"), ("tuning_angles", "tuning_angles"),
                    ("sigma_distances", "sigma_distances"), ("sigma_angles", "sigma_angles")):
        assert np.array_equal(getattr(fov, attr), g[f"fov_{name}_{k}"]), k
    assert np.allclose(fov.cell_fr_norm, g[f"fov_{name}_cell_fr_norm"], rtol=1e-14)
    got = fov.get_state(evaluate_at=None, pos=P, head_direction=HD)
    assert_rates_close(got, g[f"fov_{name}"], 2.0, f"fov {name}", pure_rel=False)
    Ag.pos, Ag.head_direction = P[0], HD[0] / np.linalg.norm(HD[0])
    assert_rates_close(fov.get_state()[:, 0], g[f"fov_{name}_agent"], 2.0, f"fov agent {name}", pure_rel=False)
    # the update() path uses the Agent's (post-motion) head direction
    Ag.update(); fov.update()
    import riab_oracle as O
    env = O.OracleEnvironment(walls=g[f"fov_{name}_walls"][4:])
    ref = O.bvc_get_state(env, fov.tuning_distances, fov.tuning_angles, fov.sigma_distances, fov.sigma_angles,
                          Ag.pos, O.TapeRNG(), min_fr=0.0, max_fr=2.0, head_direction=Ag.head_direction)[:, 0]
    assert np.abs(fov.firingrate - ref).max() <= 2e-5


@pytest.mark.parametrize("name,walls", [("open", []), ("wall", [[[0.5, 0.2], [0.5, 0.8]]])])
def test_periodic_boundary_conditions_golden(golden, name, walls):
    """Periodic box (Environment.py:130-136, :670-675, :877-879) against the live reference: teacher-forced
    steps that cross the boundary (positions wrap, measured velocity / distance take the short way round)
    and PlaceCells with wrapped distances."""
    import ratinabox_b200 as rb
    g = golden("periodic.npz")
    E = rb.Environment({"boundary_conditions": "periodic"})
    for w in walls:
        E.add_wall(w)
    assert np.array_equal(E.walls, g[f"{name}_walls"].reshape(-1, 2, 2))
    A = len(g[f"{name}_A_pos0"])
    Ag = rb.Agent(E, {"dt": 0.05, "speed_mean": 0.5, "n_agents": A})
    v0 = g[f"{name}_A_vel0"]
    Ag.pos, Ag.velocity, Ag.measured_velocity = g[f"{name}_A_pos0"], v0, v0
    Ag.rotational_velocity = np.zeros(A)
    Ag.head_direction = v0 / np.linalg.norm(v0, axis=1, keepdims=True)
    Ag.distance_travelled = np.zeros(A)
    PCs = rb.PlaceCells(Ag, {"place_cell_centres": g[f"{name}_centres"], "widths": 0.15})
    assert PCs.wall_geometry == "euclidean"
    assert_rates_close(PCs.get_state(evaluate_at=None, pos=g[f"{name}_A_pos0"]), g[f"{name}_A_pc"], 1.0, f"periodic pc {name}")
    Ag.update(_xi=g[f"{name}_A_xi"])
    assert np.abs(Ag.pos - g[f"{name}_A_pos"]).max() <= 1e-12
    assert np.abs(Ag.measured_velocity - g[f"{name}_A_mv"]).max() <= 1e-10
    assert np.abs(Ag.distance_travelled - g[f"{name}_A_dist"]).max() <= 1e-12
    assert (np.abs(g[f"{name}_A_pos"] - g[f"{name}_A_pos0"]) > 0.5).any(axis=1).sum() > 20      # wraps are exercised
    PCs.update()
    import riab_oracle as O
    env = O.OracleEnvironment(walls=walls, boundary_conditions="periodic")
    ref = O.place_cells_get_state(env, PCs.place_cell_centres, PCs.place_cell_widths, Ag.pos, O.TapeRNG()).T
    assert np.abs(PCs.firingrate - ref).max() <= 1e-5


@pytest.mark.parametrize("name,params", [
    ("lroom", {"boundary": [[0, 0], [1, 0], [1, 0.5], [0.5, 0.5], [0.5, 1], [0, 1]], "walls": [[[0.25, 0.0], [0.25, 0.3]]]}),
    ("holed", {"holes": [[[0.4, 0.4], [0.6, 0.4], [0.6, 0.6], [0.4, 0.6]]], "walls": [[[0.8, 0.0], [0.8, 0.35]]]})])
def test_polygon_boundary_and_holes_golden(golden, name, params):
    """Polygon boundary / holes against the live reference (tests/golden/polygon.npz): teacher-forced steps started
    next to boundary, hole and inner walls (repulsion + bounces off every kind of wall), line_of_sight PlaceCells with
    the reference's hard-coded `walls[4:]`, BVCs over all walls; then a free run that must stay in the environment."""
    import ratinabox_b200 as rb
    g = golden("polygon.npz")
    E = rb.Environment(dict(params))
    assert np.array_equal(E.walls, g[f"{name}_walls"]) and np.array_equal(E.extent, g[f"{name}_extent"])
    A = len(g[f"{name}_A_pos0"])
    Ag = rb.Agent(E, {"dt": 0.02, "speed_mean": 0.25, "n_agents": A})
    v0 = g[f"{name}_A_vel0"]
    Ag.pos, Ag.velocity, Ag.measured_velocity = g[f"{name}_A_pos0"], v0, v0
    Ag.rotational_velocity = np.zeros(A)
    Ag.head_direction = v0 / np.linalg.norm(v0, axis=1, keepdims=True)
    PCs = rb.PlaceCells(Ag, {"place_cell_centres": g[f"{name}_centres"], "widths": 0.15})
    assert PCs.wall_geometry == "line_of_sight"
    td, ta, sd, sa = g[f"{name}_bvc"]
    BVCs = rb.BoundaryVectorCells(Ag, {"tuning_distance": td, "tuning_angle": np.degrees(ta), "sigma_distance": sd,
                                       "sigma_angle": np.degrees(sa)})
    assert_rates_close(PCs.get_state(evaluate_at=None, pos=g[f"{name}_A_pos0"]), g[f"{name}_A_pc"], 1.0, f"polygon pc {name}")
    assert_rates_close(BVCs.get_state(evaluate_at=None, pos=g[f"{name}_A_pos0"]), g[f"{name}_A_bvc"], 1.0, f"polygon bvc {name}")
    Ag.update(_xi=g[f"{name}_A_xi"])
    assert np.abs(Ag.pos - g[f"{name}_A_pos"]).max() <= 1e-12
    assert np.abs(Ag.velocity - g[f"{name}_A_vel"]).max() <= 1e-12
    assert np.abs(Ag.measured_velocity - g[f"{name}_A_mv"]).max() <= 1e-10
    Ag.run(300)                                            # Philox-driven: every agent stays strictly inside
    pos = Ag.pos
    assert all(E.check_if_position_is_in_environment(p) for p in pos)
    hp = Ag.get_history_arrays()["pos"]
    assert np.isfinite(hp).all()


OVC_WALLS = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]
OVC_OBJECTS = [([0.15, 0.2], 0), ([0.5, 0.8], "same"), ([0.85, 0.3], "new"), ([0.5, 0.25], "new"), ([0.9, 0.9], 1)]


def _ovc_setup(rb, g, A):
    E = rb.Environment()
    for w in OVC_WALLS:
        E.add_wall(w)
    for o, t in OVC_OBJECTS:
        E.add_object(o, type=t)
    assert np.array_equal(E.objects["objects"], g["objects"]) and np.array_equal(E.objects["object_types"], g["object_types"])
    Ag = rb.Agent(E, {"dt": 0.02, "n_agents": A, "seed": 3})
    pops = {}
    for k, cls, extra in (("allo", rb.ObjectVectorCells, {}), ("eucl", rb.ObjectVectorCells, {"walls_occlude": False}),
                          ("fov", rb.FieldOfViewOVCs, {"spatial_resolution": 0.05})):
        td, ta, sd, sa = g[f"{k}_tuning"]
        P = cls(Ag, dict({"object_tuning_type": [int(t) for t in g[f"{k}_types"]], "tuning_distance": td,
                          "tuning_angle": np.degrees(ta), "sigma_distance": sd, "sigma_angle": np.degrees(sa),
                          "cell_arrangement": "random"}, **extra))
        # the manifold's tuning is pinned to the fixture's (field-of-view cells come from diverging_manifold there)
        P.tuning_distances, P.tuning_angles, P.sigma_distances, P.sigma_angles = td.copy(), ta.copy(), sd.copy(), sa.copy()
        assert P.n == len(td) and P.wall_geometry == str(g[f"{k}_geom"])
        pops[k] = P
    assert pops["fov"].reference_frame == "egocentric"
    return E, Ag, pops


def test_object_vector_cells_golden(golden):
    """ObjectVectorCells (allocentric, occluding walls / euclidean) and FieldOfViewOVCs (egocentric) against the live
    reference at 384 positions / head directions (tests/golden/ovc.npz)."""
    import ratinabox_b200 as rb
    g = golden("ovc.npz")
    E, Ag, pops = _ovc_setup(rb, g, 4)
    for k in ("allo", "eucl"):
        assert_rates_close(pops[k].get_state(evaluate_at=None, pos=g["A_pos"]), g[f"A_{k}"], 1.0, f"ovc {k}")
    assert_rates_close(pops["fov"].get_state(evaluate_at=None, pos=g["A_pos"], head_direction=g["A_hd"]), g["A_fov"], 1.0, "ovc fov")
    one = pops["fov"].get_state(evaluate_at=None, pos=g["A_pos"][:5], head_direction=g["A_hd"][2])   # one direction for all
    import riab_oracle as O
    env = O.OracleEnvironment(walls=OVC_WALLS)
    td, ta, sd, sa = g["fov_tuning"]
    ref = O.ovc_get_state(env, g["objects"], g["object_types"], td, ta, sd, sa, g["fov_types"], g["A_pos"][:5], O.TapeRNG(),
                          "line_of_sight", head_direction=g["A_hd"][2])
    assert np.abs(one - ref).max() <= 1e-5


def test_object_vector_cells_stepped_with_the_agent(golden):
    """Agent.update + three ObjectVectorCells populations: the fused step (first population), riab_neurons_update
    (the others) and riab_run all see the positions / head directions of the same step; spikes and history rows."""
    import ratinabox_b200 as rb
    import riab_oracle as O
    g = golden("ovc.npz")
    A = 96
    E, Ag, pops = _ovc_setup(rb, g, A)
    env = O.OracleEnvironment(walls=OVC_WALLS)

    def check(tag):
        pos, hd = Ag.pos, Ag.head_direction
        for k, ego in (("allo", False), ("eucl", False), ("fov", True)):
            td, ta, sd, sa = g[f"{k}_tuning"]
            ref = O.ovc_get_state(env, g["objects"], g["object_types"], td, ta, sd, sa, g[f"{k}_types"], pos, O.TapeRNG(),
                                  str(g[f"{k}_geom"]), head_direction=(hd if ego else None)).T
            assert np.abs(pops[k].firingrate - ref).max() <= 1e-5, (tag, k)

    for _ in range(3):
        Ag.update()
        for P in pops.values():
            P.update()
    check("stepped")
    Ag.run(5)
    check("run")
    for k, P in pops.items():
        h = P.get_history_arrays()
        assert h["firingrate"].shape == (8, A, P.n) and h["spikes"].shape == (8, A, P.n)
        assert np.array_equal(h["firingrate"][-1], P.firingrate)


def _param_variants():
    import test_oracle_golden as T
    return T.PARAM_VARIANTS


@pytest.mark.parametrize("name", sorted(_param_variants()))
def test_agent_parameter_and_keyword_variants_golden(golden, name):
    """The CUDA motion step against the live reference for every motion parameter / Agent.update keyword
    (tests/golden/modeA_params.npz; same table as the oracle's test): positions and velocities to 1e-12."""
    import ratinabox_b200 as rb
    g = golden("modeA_params.npz")
    params, kw = _param_variants()[name]
    kw = dict(kw)
    ratio = kw.pop("drift_ratio", None)
    A = len(g["pos0"])
    E = rb.Environment()
    for w in OVC_WALLS:                                # the two-wall box of config 2
        E.add_wall(w)
    assert np.array_equal(E.walls, g["walls"])
    Ag = rb.Agent(E, dict({"dt": 0.01, "n_agents": A}, **params))
    Ag.pos, Ag.velocity, Ag.measured_velocity = g["pos0"], g["vel0"], g["mv0"]
    Ag.rotational_velocity, Ag.head_direction = g["rot0"], g["hd0"]
    Ag.distance_travelled = np.zeros(A)
    if ratio is not None:
        Ag.update(drift_velocity=g["drift"], drift_to_random_strength_ratio=ratio, _xi=g["xi"], **kw)
    else:
        Ag.update(_xi=g["xi"], **kw)
    tol = {"pos": 1e-12, "vel": 1e-12, "rot": 1e-10, "mv": 1e-9, "mrot": 1e-6, "hd": 1e-10, "dist": 1e-12, "dclose": 1e-12}
    got = {"pos": Ag.pos, "vel": Ag.velocity, "rot": Ag.rotational_velocity, "mv": Ag.measured_velocity,
           "mrot": Ag.measured_rotational_velocity, "hd": Ag.head_direction, "dist": Ag.distance_travelled,
           "dclose": Ag.distance_to_closest_wall}
    for key in tol:
        ref = g[f"{name}_{key}"]
        if key == "dclose" and name == "no_repel":
            continue                                   # Agent.py:359-360 returns before distance_to_closest_wall is set
        err = np.abs(np.asarray(got[key]).reshape(ref.shape) - ref).max()
        assert err <= tol[key], (name, key, err)
    if name == "dt_arg":
        assert Ag.dt == 0.05
