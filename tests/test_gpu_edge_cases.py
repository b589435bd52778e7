"""Edge cases the reference's usage implies: ragged sizes (agents not a multiple of the 32-agent tile,
cells not a multiple of 4 / 128, more cells than one consumer pass holds), every register-template of the
line-of-sight kernels (1, 2, 3..4, 5..8 inner walls), odd shard offsets, noise, populations sharing an
Agent.  GPU only."""
import numpy as np
import pytest

import riab_oracle as O
from philox_np import expected_spikes, expected_spikes_of

pytestmark = pytest.mark.gpu


def _walls(k):
    """k internal walls, alternately from floor and ceiling (none touches another)."""
    out = []
    for i in range(k):
        x = (i + 1) / (k + 1)
        out.append([[x, 0.0], [x, 0.45]] if i % 2 == 0 else [[x, 1.0], [x, 0.55]])
    return out


def _make(rb, A, walls, seed=4, **agent):
    np.random.seed(seed)
    E = rb.Environment()
    for w in walls:
        E.add_wall(w)
    Ag = rb.Agent(E, dict({"dt": 0.01, "n_agents": A, "seed": 9}, **agent))
    return E, Ag


@pytest.mark.parametrize("A,N,k", [(1, 1, 0), (3, 5, 1), (33, 130, 2), (100, 1025, 3), (70, 2500, 0), (31, 257, 7), (64, 4096, 2)])
def test_ragged_sizes_all_wall_templates(A, N, k):
    import ratinabox_b200 as rb
    walls = _walls(k)
    E, Ag = _make(rb, A, walls)
    geom = "line_of_sight" if k else "euclidean"
    PCs = rb.PlaceCells(Ag, {"n": N, "wall_geometry": geom, "widths": 0.15, "max_fr": 4.0})
    GCs = rb.GridCells(Ag, {"n": max(N // 3, 1)})
    assert PCs.n == N
    steps = 3
    pos0 = np.array(Ag.pos, dtype=float).reshape(A, 2).copy()
    Ag.run(steps)
    pos = np.array(Ag.pos, dtype=float).reshape(A, 2)
    env = O.OracleEnvironment(walls=walls)
    ref = O.place_cells_get_state(env, PCs.place_cell_centres, PCs.place_cell_widths, pos, O.TapeRNG(),
                                  "gaussian", geom, 0.0, 4.0).T
    fr = np.array(PCs.firingrate, dtype=float).reshape(A, N)
    assert np.abs(fr - ref).max() <= 4e-5            # 1e-5 of the rate scale (max_fr = 4)
    refg = O.grid_cells_get_state(GCs.gridscales, GCs.phase_offsets, GCs.w, pos).T
    assert np.abs(np.array(GCs.firingrate, dtype=float).reshape(A, GCs.n) - refg).max() <= 1e-5
    h = PCs.get_history_arrays()
    want = (steps, N) if A == 1 else (steps, A, N)
    assert h["firingrate"].shape == want and h["spikes"].shape == want
    # spikes of the last step against the NumPy mirror of the Philox stream (population 0)
    sp = h["spikes"][-1].reshape(A, N)
    assert np.array_equal(sp, expected_spikes_of(PCs, 9, steps - 1, np.arange(A), fr, 0.01, pop=0, fr_bound=4.0))
    assert np.isfinite(pos).all() and not np.array_equal(pos, pos0)


def test_odd_shard_offsets_give_identical_results():
    """Shards that start at odd global ids use the unpaired spike path and the same Philox streams."""
    import ratinabox_b200 as rb
    A = 77
    E, Ag = _make(rb, A, _walls(2))
    PCs = rb.PlaceCells(Ag, {"n": 96})
    pos0, vel0 = Ag.pos.copy(), Ag.velocity.copy()
    Ag.run(5)
    full_pos, full_fr, full_sp = Ag.pos, PCs.firingrate, PCs.get_history_arrays()["spikes"]
    for start, stop in ((0, 33), (33, 77)):
        E2, Ag2 = _make(rb, stop - start, _walls(2), id_offset=start)
        Ag2.pos, Ag2.velocity, Ag2.measured_velocity = pos0[start:stop], vel0[start:stop], vel0[start:stop]
        P2 = rb.PlaceCells(Ag2, {"place_cell_centres": PCs.place_cell_centres})
        Ag2.run(5)
        assert np.array_equal(Ag2.pos, full_pos[start:stop])
        assert np.array_equal(P2.firingrate, full_fr[start:stop])
        assert np.array_equal(P2.get_history_arrays()["spikes"], full_sp[:, start:stop])


def test_noise_is_ornstein_uhlenbeck_with_the_requested_std():
    """Neurons.update's OU noise (Neurons.py:153-160): stationary std = noise_std, coherence noise_coherence_time."""
    import ratinabox_b200 as rb
    E, Ag = _make(rb, 64, [])
    PCs = rb.PlaceCells(Ag, {"n": 64, "noise_std": 0.3, "noise_coherence_time": 0.05, "wall_geometry": "euclidean"})
    clean = rb.PlaceCells(Ag, {"place_cell_centres": PCs.place_cell_centres, "wall_geometry": "euclidean"})
    resid = []
    for s in range(400):
        Ag.update(); PCs.update(); clean.update()
        if s >= 100:
            resid.append(PCs.firingrate - clean.firingrate)
    r = np.array(resid)
    assert abs(r.mean()) < 0.01 and abs(r.std() - 0.3) < 0.02
    lag = 5                                       # 5 steps = one coherence time -> correlation exp(-1)
    c = np.mean(r[lag:] * r[:-lag]) / r.var()
    assert abs(c - np.exp(-1.0)) < 0.06


def test_three_populations_share_one_agent_in_run():
    """config 5 in miniature: Place + Grid + BVC populations of one Agent stepped by riab_run all see the
    positions of the same step (population 0 skewed, the others rate-only)."""
    import ratinabox_b200 as rb
    walls = _walls(2)
    E, Ag = _make(rb, 48, walls)
    PCs = rb.PlaceCells(Ag, {"n": 64})
    GCs = rb.GridCells(Ag, {"n": 30})
    BVCs = rb.BoundaryVectorCells(Ag, {"n": 20})
    Ag.run(6)
    pos = Ag.pos
    env = O.OracleEnvironment(walls=walls)
    rng = O.TapeRNG()
    assert np.abs(PCs.firingrate - O.place_cells_get_state(env, PCs.place_cell_centres, PCs.place_cell_widths, pos, rng,
                                                           "gaussian", "line_of_sight").T).max() <= 1e-5
    assert np.abs(GCs.firingrate - O.grid_cells_get_state(GCs.gridscales, GCs.phase_offsets, GCs.w, pos).T).max() <= 1e-5
    assert np.abs(BVCs.firingrate - O.bvc_get_state(env, BVCs.tuning_distances, BVCs.tuning_angles, BVCs.sigma_distances,
                                                    BVCs.sigma_angles, pos, rng).T).max() <= 1e-5
    hp = Ag.get_history_arrays()["pos"]
    assert hp.shape == (6, 48, 2) and np.abs(hp[-1] - pos).max() <= 1e-6
    # history rows of every population belong to the same 6 steps
    for ns in (PCs, GCs, BVCs):
        assert ns.get_history_arrays()["firingrate"].shape[0] == 6


@pytest.mark.parametrize("widths", ["uniform", "mixed"])
def test_line_of_sight_on_the_decision_boundary(widths):
    """Adversarial geometry for the float32 line-of-sight predicate: positions chosen so that the segment
    centre -> position grazes a wall END POINT (l_b = 0 or 1 up to rounding), runs ALONG the wall's line, or starts
    on it.  Those pairs fall inside the float32 band and must be decided by the exact float64 path exactly like
    utils.vector_intercepts (utils.py:96-106): a single misclassified pair changes a rate by O(1).
    `uniform` widths take the expanded-exponent kernel, `mixed` the direct one."""
    import ratinabox_b200 as rb
    rng = np.random.default_rng(3)
    walls = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]]]
    N = 256
    centres = rng.uniform(0.02, 0.98, size=(N, 2))
    centres[:8, 0] = 0.3                               # centres ON the first wall's line (above and on the wall)
    centres[8:12, 0] = 0.7
    w = np.full(N, 0.2) if widths == "uniform" else rng.uniform(0.1, 0.3, size=N)
    ends = np.array([[0.3, 0.5], [0.7, 0.5]])
    pos = []
    for k in range(1500):                              # through a wall end: p = e + s (e - c)
        c, e = centres[rng.integers(N)], ends[rng.integers(2)]
        s = rng.uniform(0.05, 3.0)
        pos.append(e + s * (e - c))
    for k in range(250):                               # on a wall's line (blocked-ness of centres on the other line)
        pos.append([0.3 if k % 2 else 0.7, rng.uniform(0.02, 0.98)])
    pos = np.array(pos)
    pos = pos[(pos[:, 0] > 0.01) & (pos[:, 0] < 0.99) & (pos[:, 1] > 0.01) & (pos[:, 1] < 0.99)]
    assert len(pos) > 600
    np.random.seed(1)
    E = rb.Environment()
    for wl in walls:
        E.add_wall(wl)
    Ag = rb.Agent(E, {"dt": 0.01, "n_agents": 4})
    PCs = rb.PlaceCells(Ag, {"place_cell_centres": centres, "widths": w if widths == "mixed" else 0.2,
                             "wall_geometry": "line_of_sight"})
    got = PCs.get_state(evaluate_at=None, pos=pos)     # (N, n_pos)
    env = O.OracleEnvironment(walls=walls)
    ref = O.place_cells_get_state(env, centres, w, pos, O.TapeRNG(), "gaussian", "line_of_sight")
    dist = O.distances_accounting_for_environment(env, centres, pos, "line_of_sight", O.TapeRNG())
    assert 0.05 < (dist >= 1000).mean() < 0.8          # both outcomes are well represented
    bad = np.abs(got - ref) > 1e-5
    assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:5], got[bad][:5], ref[bad][:5])


def test_more_rows_than_a_grid_dimension():
    """70 000 agents (> 65 535, the y/z grid limit): every per-row kernel (BVC rays / integration / spike post-pass,
    one_hot) must index rows on the x dimension."""
    import ratinabox_b200 as rb
    A = 70000
    E, Ag = _make(rb, A, _walls(1))
    BVCs = rb.BoundaryVectorCells(Ag, {"n": 8})
    OH = rb.PlaceCells(Ag, {"n": 16, "description": "one_hot", "wall_geometry": "euclidean"})
    Ag.update(); BVCs.update(); OH.update()
    pos = Ag.pos
    sel = np.array([0, 1, 65534, 65535, 65536, A - 1])
    env = O.OracleEnvironment(walls=_walls(1))
    ref = O.bvc_get_state(env, BVCs.tuning_distances, BVCs.tuning_angles, BVCs.sigma_distances, BVCs.sigma_angles,
                          pos[sel], O.TapeRNG()).T
    assert np.abs(BVCs.firingrate[sel] - ref).max() <= 1e-5
    assert np.array_equal(OH.firingrate.sum(axis=1), np.ones(A))
    assert BVCs.get_history_arrays()["spikes"].shape == (1, A, 8)


def test_population_parameters_beyond_the_defaults():
    """Neurons-level parameters the fixtures keep at their defaults: BVC dtheta (T = 72 and 360 test angles),
    GridCells width_ratio / shifted_cosines and PlaceCells profiles with [min_fr, max_fr] != [0, 1]."""
    import ratinabox_b200 as rb
    walls = _walls(3)
    E, Ag = _make(rb, 40, walls)
    pos = np.random.RandomState(2).uniform(0.03, 0.97, size=(300, 2))
    env = O.OracleEnvironment(walls=walls)
    rng = O.TapeRNG()
    for dtheta in (5, 1):
        B = rb.BoundaryVectorCells(Ag, {"n": 24, "dtheta": dtheta, "min_fr": 0.5, "max_fr": 3.0})
        assert len(B.test_angles) == 360 // dtheta
        ref = O.bvc_get_state(env, B.tuning_distances, B.tuning_angles, B.sigma_distances, B.sigma_angles, pos, rng,
                              dtheta=dtheta, min_fr=0.5, max_fr=3.0)
        assert np.abs(B.get_state(evaluate_at=None, pos=pos) - ref).max() <= 2.5e-5      # 1e-5 of the 2.5 Hz span
    for desc in ("rectified_cosines", "shifted_cosines"):
        G = rb.GridCells(Ag, {"n": 30, "description": desc, "width_ratio": 0.5, "min_fr": -1.0, "max_fr": 2.0})
        ref = O.grid_cells_get_state(G.gridscales, G.phase_offsets, G.w, pos, desc, 0.5, -1.0, 2.0)
        assert np.abs(G.get_state(evaluate_at=None, pos=pos) - ref).max() <= 3e-5
    for desc in ("gaussian_threshold", "diff_of_gaussians", "top_hat"):
        P = rb.PlaceCells(Ag, {"n": 40, "description": desc, "widths": 0.25, "min_fr": 1.0, "max_fr": 5.0,
                               "wall_geometry": "line_of_sight"})
        ref = O.place_cells_get_state(env, P.place_cell_centres, P.place_cell_widths, pos, rng, desc, "line_of_sight",
                                      1.0, 5.0, scalar_width=0.25)
        err = np.abs(P.get_state(evaluate_at=None, pos=pos) - ref)
        assert err.max() <= 4e-5, (desc, err.max())


def test_nan_position_gives_zero_rates_like_the_reference():
    """Neurons.update: `if np.isnan(self.Agent.pos[0]): firingrate = zeros` (ratinabox/Neurons.py:163-164) -- per agent here,
    for every cell type, in the stepped API and in get_state-free updates; the other agents are unaffected."""
    import ratinabox_b200 as rb
    A = 70
    E, Ag = _make(rb, A, _walls(2))
    PCs = rb.PlaceCells(Ag, {"n": 96, "min_fr": 0.2, "max_fr": 3.0})
    GCs = rb.GridCells(Ag, {"n": 40})
    BVCs = rb.BoundaryVectorCells(Ag, {"n": 24})
    NZ = rb.PlaceCells(Ag, {"n": 32, "noise_std": 0.1})
    pos = Ag.pos.copy()
    bad = [0, 33, 69]
    pos[bad, 0] = np.nan
    Ag.pos = pos
    for ns in (PCs, GCs, BVCs, NZ):
        ns.update()                                   # rates at the current positions (no motion step queued)
    for ns in (PCs, GCs, BVCs):
        fr = ns.firingrate
        assert np.array_equal(fr[bad], np.zeros((len(bad), ns.n))), type(ns).__name__
        good = np.setdiff1d(np.arange(A), bad)
        assert np.isfinite(fr[good]).all() and np.abs(fr[good]).max() > 0
        assert not ns.get_history_arrays()["spikes"][-1][bad].any()
    frz = NZ.firingrate                               # zeros + OU noise (Neurons.py:167-168)
    assert np.isfinite(frz[bad]).all() and 0 < np.abs(frz[bad]).max() < 1.0


def test_zero_length_wall_is_a_point_obstacle():
    """The reference's own test adds a zero-length wall (tests/test_environment.py:20-23).  Its jittered arithmetic treats it
    as a ~1e-6 m segment; the engine's zero-jitter arithmetic treats it as the point itself instead of dividing 0 / 0
    (utils.py:121-184): positions stay finite and equal the oracle's for the same box with a 1e-9 m wall at that point."""
    import ratinabox_b200 as rb
    A = 300
    pt = [0.4, 0.4]
    E, Ag = _make(rb, A, [[pt, pt]])
    assert len(E.walls) == 5
    rs = np.random.RandomState(3)
    pos0 = np.clip(np.array(pt) + rs.normal(0, 0.06, (A, 2)), 0.02, 0.98)      # many agents inside the repulsion radius
    Ag.pos = pos0
    vel0 = Ag.velocity.copy()
    xi = rs.normal(size=(A, 2))
    PCs = rb.PlaceCells(Ag, {"n": 64, "wall_geometry": "line_of_sight"})
    BVCs = rb.BoundaryVectorCells(Ag, {"n": 16})
    Ag.update(_xi=xi)
    PCs.update(); BVCs.update()
    pos1 = Ag.pos
    assert np.isfinite(pos1).all() and np.isfinite(PCs.firingrate).all() and np.isfinite(BVCs.firingrate).all()
    env = O.OracleEnvironment(walls=[[pt, [pt[0] + 1e-9, pt[1]]]])
    near = 0
    for a in range(A):
        oa = O.OracleAgent(env, pos0[a], vel0[a], {"dt": 0.01})
        oa.update(O.TapeRNG(agent_xi=xi[a]))
        assert np.abs(oa.pos - pos1[a]).max() <= 1e-6, a
        near += np.linalg.norm(pos0[a] - pt) < 0.1
    assert near > 50
    for _ in range(200):
        Ag.update()
    assert np.isfinite(Ag.pos).all()


def test_agent_exactly_on_a_wall_stays_finite():
    """`Ag.pos = [0.5, 0.5]` with a wall through x = 0.5: the reference's 1e-6 jitter hides the 0 / 0 of the unit normal
    (Agent.py:370, utils.py:143-144); the engine skips that wall's repulsion for the step instead of poisoning the state."""
    import ratinabox_b200 as rb
    E, Ag = _make(rb, 8, [[[0.5, 0.2], [0.5, 0.8]]])
    pos = Ag.pos.copy()
    pos[:4] = [0.5, 0.5]
    Ag.pos = pos
    PCs = rb.PlaceCells(Ag, {"n": 32})
    for _ in range(50):
        Ag.update(); PCs.update()
    assert np.isfinite(Ag.pos).all() and np.isfinite(Ag.velocity).all() and np.isfinite(PCs.firingrate).all()


@pytest.mark.parametrize("scale,n_extra", [(1.0, 0), (10.0, 0), (1.0, -6), (1.0, 10), (1.0, 34)])
def test_bvc_first_wall_exact_under_the_float32_screen(scale, n_extra):
    """k_bvc_rays drops walls with a float32 screen (l_b margin, l_a certainly negative / certainly behind another wall)
    before the float64 walk: distance to the first wall (as float32) and the wall index equal the oracle's argmax
    (Neurons.py:1651-1684) for positions on wall lines, at wall ends and corners, 1e-9 next to walls, and random ones."""
    import ctypes as C
    import torch
    import ratinabox_b200 as rb
    from ratinabox_b200 import _lib
    walls = [[[0.3, 0.0], [0.3, 0.5]], [[0.7, 1.0], [0.7, 0.5]], [[0.3, 0.5], [0.5, 0.7]], [[0.1, 0.8], [0.45, 0.8]],
             [[0.6, 0.2], [0.9, 0.2]], [[0.9, 0.2], [0.9, 0.45]]]
    # n_extra: -6 = the empty box (4 walls); > 0 = short random extra walls: 20 walls take the generic screen with a 32-bit
    # mask (the table screen holds 16), 44 walls the 64-bit one
    rsw = np.random.RandomState(17)
    if n_extra < 0:
        walls = []
    for _ in range(max(n_extra, 0)):
        a = rsw.uniform(0.05, 0.95, 2)
        th = rsw.uniform(0, np.pi)
        b = np.clip(a + rsw.uniform(0.03, 0.15) * np.array([np.cos(th), np.sin(th)]), 0.02, 0.98)
        walls.append([a.tolist(), b.tolist()])
    walls = (np.array(walls).reshape(-1, 2, 2) * scale).tolist()
    E = rb.Environment({"scale": scale})
    for w in walls:
        E.add_wall(w)
    Ag = rb.Agent(E, {"dt": 0.01})
    bvc = rb.BoundaryVectorCells(Ag, {"n": 8})
    rs = np.random.RandomState(3)
    P = [rs.uniform(0, scale, (300, 2))]
    aw = np.array(E.walls, dtype=float)
    for w in aw:                                        # on the wall line, at its ends, just off it, on its extension
        t = rs.uniform(-0.3, 1.3, 12)[:, None]
        on = w[0] + t * (w[1] - w[0])
        nrm = np.array([-(w[1] - w[0])[1], (w[1] - w[0])[0]]) / np.linalg.norm(w[1] - w[0])
        P += [on, on + 1e-9 * scale * nrm, on - 1e-7 * scale * nrm, w[0][None], w[1][None], w[0][None] + 1e-12]
    P = np.clip(np.concatenate(P), 0.0, scale)
    n, T = len(P), bvc.n_test_angles
    pos = torch.as_tensor(P, device=bvc.device)
    out = torch.empty((n, 8), dtype=torch.float32, device=bvc.device)
    first = torch.full((n, T), -7, dtype=torch.int32, device=bvc.device)
    bvc._rates_from_positions(pos, n, out, first_wall=first)
    torch.cuda.synchronize()
    nt = (n + 31) // 32
    d_gpu = bvc._scratch_for(n)[:nt * T * 32].reshape(nt, T, 32).permute(0, 2, 1).reshape(nt * 32, T)[:n].cpu().numpy()
    env = O.OracleEnvironment(scale=scale, walls=walls)
    d, fw = O.bvc_first_wall_distances(env, P, np.asarray(bvc.test_directions), O.TapeRNG())
    assert np.array_equal(first.cpu().numpy(), fw.astype(np.int32))
    with np.errstate(over="ignore"):
        assert np.array_equal(d_gpu, d.astype(np.float32), equal_nan=True)


@pytest.mark.parametrize("n,lo,hi", [(100, 4.0, 9.0), (64, 11.25, 11.25), (200, 6.0, 30.0), (31, 3.0, 4.0)])
def test_bvc_angular_windows_skip_only_negligible_terms(n, lo, hi):
    """riab_bvc_pack sorts the cells by tuning angle and k_bvc_integrate skips, per warp of 32 slots, the test angles where
    every von Mises weight is < 2^-30 of its peak: rates of narrowly tuned populations (windows much shorter than the
    circle) equal the oracle's full sums (Neurons.py:1710-1744) to 1e-5 of the rate scale, in the cells' own order."""
    import ratinabox_b200 as rb
    walls = _walls(3)
    E, Ag = _make(rb, 1, walls)
    rs = np.random.RandomState(n)
    bvc = rb.BoundaryVectorCells(Ag, {"n": n, "min_fr": 0.0, "max_fr": 3.0})
    bvc.tuning_angles = rs.uniform(0, 2 * np.pi, n)
    bvc.sigma_angles = np.radians(rs.uniform(lo, hi, n))
    bvc.tuning_distances = rs.uniform(0.05, 0.5, n)
    bvc.sigma_distances = bvc.tuning_distances / 12 + 0.08
    P = rs.uniform(0.02, 0.98, (150, 2))
    got = bvc.get_state(evaluate_at=None, pos=P)
    env = O.OracleEnvironment(walls=walls)
    ref = O.bvc_get_state(env, bvc.tuning_distances, bvc.tuning_angles, bvc.sigma_distances, bvc.sigma_angles, P, O.TapeRNG(),
                          min_fr=0.0, max_fr=3.0)
    assert got.shape == ref.shape == (n, 150)
    assert np.abs(got - ref).max() <= 3e-5, np.abs(got - ref).max()
    big = ref > 3e-3
    assert (np.abs(got - ref)[big] / ref[big]).max() <= 2e-5
