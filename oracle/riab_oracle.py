"""TEST INFRASTRUCTURE -- CPU oracle for the RatInABox per-step hot path.

This file is a float64 NumPy restatement of the reference algorithm
(RatInABox v1.15.3, /root/reference/ratinabox) for the path named in
BASELINE.json: ``Agent.update`` (2D, solid rectangular box + internal walls)
and ``Neurons.update`` / ``get_state`` for PlaceCells, GridCells and
allocentric BoundaryVectorCells.  Every function cites the reference
file:line it follows.

It is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  ``ratinabox_b200`` never does.

Parity pin: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the live,
unmodified reference run in the build container (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).  With
``GlobalRNG`` the oracle consumes NumPy's global legacy RNG in exactly the
reference's order (SURVEY.md section 8a "RNG tape"), so under the same
``np.random.seed`` it reproduces the reference bit for bit, jitter included.
With ``TapeRNG`` (zero jitter, injected standard normals) it is the
deterministic configuration the CUDA path is compared against.

Arithmetic in third-party dependencies that are not under /root/reference:
``scipy.stats.norm.ppf/cdf`` (utils.py:411,420) are ``scipy.special.ndtri`` /
``ndtr`` (SciPy 1.18.1 here; unpinned in the reference's setup.cfg:24-28) and
are called as such.
"""
from __future__ import annotations

import numpy as np
from scipy import special as _sp


# --------------------------------------------------------------------------- RNG
class GlobalRNG:
    """Draw from numpy's global legacy RNG with the reference's call shapes."""

    def jitter(self, scale, shape):
        return np.random.normal(scale=scale, size=shape)

    def ou_normal(self, dt, shape):
        # utils.py:367  np.random.normal(size=x.shape, scale=dt)
        return np.random.normal(size=shape, scale=dt)

    def uniform(self, shape):
        return np.random.uniform(0, 1, size=shape)

    def randn(self, n):
        return np.random.randn(n)


class TapeRNG:
    """Mode A of SURVEY.md section 8(c): geometry jitter is zero, the OU draws are
    popped from supplied standard-normal tapes (value = dt * xi, which is
    exactly what ``np.random.normal(scale=dt)`` computes from its gaussian)."""

    def __init__(self, agent_xi=None, noise_xi=None, spike_u=None):
        self.agent_xi = list(np.asarray(agent_xi, dtype=float).reshape(-1)) if agent_xi is not None else []
        self.noise_xi = noise_xi  # iterable of (N,) arrays or None -> zeros
        self.spike_u = spike_u    # iterable of (N,) arrays or None -> ones (no spikes)
        self._noise_i = 0
        self._spike_i = 0

    def jitter(self, scale, shape):
        return np.zeros(shape)

    def ou_normal(self, dt, shape):
        """Agent OU draws (scalars): pop standard normals from the agent tape."""
        shape = tuple(shape)
        n = int(np.prod(shape)) if len(shape) else 1
        vals = [self.agent_xi.pop(0) if self.agent_xi else 0.0 for _ in range(n)]
        return dt * np.asarray(vals, dtype=float).reshape(shape)

    def uniform(self, shape):
        if self.spike_u is None:
            return np.ones(shape)
        u = np.asarray(self.spike_u[self._spike_i], dtype=float).reshape(shape)
        self._spike_i += 1
        return u

    def randn(self, n):
        return np.zeros(n)


# ---------------------------------------------------------------- geometry (utils)
def vector_intercepts(seg_a, seg_b, rng, return_collisions=False):
    """utils.py:30-118.  seg_a (Na,2,2), seg_b (Nb,2,2) -> intercept parameters
    (Na,Nb,2) = (l_a, l_b), or the strict-interior collision mask (Na,Nb)."""
    a = np.asarray(seg_a, dtype=float).reshape(-1, 2, 2)
    b = np.asarray(seg_b, dtype=float).reshape(-1, 2, 2)
    a = a + rng.jitter(1e-9, a.shape)          # utils.py:64-66
    b = b + rng.jitter(1e-9, b.shape)          # utils.py:67-69
    d0 = b[None, :, 0, :] - a[:, None, 0, :]   # (Na,Nb,2)  utils.py:74-76
    sa = (a[:, 1, :] - a[:, 0, :])[:, None, :]  # (Na,1,2)
    sb = (b[:, 1, :] - b[:, 0, :])[None, :, :]  # (1,Nb,2)
    # perpendiculars [x,y]_p = [-y,x]  (utils.py:79-82)
    sa_px, sa_py = -sa[..., 1], sa[..., 0]
    sb_px, sb_py = -sb[..., 1], sb[..., 0]
    # utils.py:96-97 (two-term sums in x,y order)
    l_a = (d0[..., 0] * sb_px + d0[..., 1] * sb_py) / (sa[..., 0] * sb_px + sa[..., 1] * sb_py)
    l_b = ((-d0[..., 0]) * sa_px + (-d0[..., 1]) * sa_py) / (sb[..., 0] * sa_px + sb[..., 1] * sa_py)
    if return_collisions:
        return (l_a > 0) & (l_a < 1) & (l_b > 0) & (l_b < 1)   # utils.py:101-106
    return np.stack((l_a, l_b), axis=-1)


def shortest_vectors_from_points_to_lines(positions, segments, rng):
    """utils.py:121-184 -> (Np,Nv,2) vectors from the segments to the points."""
    p = np.asarray(positions, dtype=float).reshape(-1, 2)
    v = np.asarray(segments, dtype=float).reshape(-1, 2, 2)
    p = p + rng.jitter(1e-6, p.shape)           # utils.py:143
    v = v + rng.jitter(1e-6, v.shape)           # utils.py:144
    d = p[:, None, :] - v[None, :, 0, :]
    s = v[:, 1, :] - v[:, 0, :]
    l_v = (d[..., 0] * s[:, 0] + d[..., 1] * s[:, 1]) / (s[:, 0] * s[:, 0] + s[:, 1] * s[:, 1])
    l_v = np.where(l_v > 1, 1.0, l_v)           # utils.py:169-170
    l_v = np.where(l_v < 0, 0.0, l_v)
    return p[:, None, :] - (v[None, :, 0, :] + l_v[..., None] * s[None, :, :])  # utils.py:182


def get_angle(vec):
    """utils.py:231-273 for a single 2-vector: note the eps added to x (utils.py:258-260)."""
    vec = np.asarray(vec, dtype=float)
    return np.mod(np.arctan2(vec[1], vec[0] + 1e-6), 2 * np.pi)


def pi_domain(x):
    """utils.py:331-341."""
    x = np.asarray(x, dtype=float) % (2 * np.pi)
    return np.where(x > np.pi, -2 * np.pi + x, x)


def rotate(vec, theta):
    """utils.py:293-301."""
    R = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    return np.matmul(R, vec)


def wall_bounce(velocity, wall):
    """utils.py:304-328: reflect ``velocity`` in ``wall`` ((2,2) endpoints)."""
    par = wall[1] - wall[0]
    perp = np.array([-par[1], par[0]])          # utils.py:17-27
    if np.dot(perp, velocity) <= 0:
        perp = -perp
    if np.dot(par, velocity) <= 0:
        par = -par
    par, perp = par / np.linalg.norm(par), perp / np.linalg.norm(perp)
    return par * np.dot(velocity, par) - perp * np.dot(velocity, perp)


def ornstein_uhlenbeck(dt, x, drift, noise_scale, coherence_time, rng):
    """utils.py:347-368.  Returns dx."""
    x = np.array(x, dtype=float)
    drift = drift * np.ones_like(x)
    noise_scale = noise_scale * np.ones_like(x)
    coherence_time = coherence_time * np.ones_like(x)
    sigma = np.sqrt((2 * noise_scale ** 2) / (coherence_time * dt))
    theta = 1 / coherence_time
    return theta * (drift - x) * dt + sigma * rng.ou_normal(dt, x.shape)


def normal_to_rayleigh(x, sigma=1.0):
    """utils.py:409-413 (scipy.stats.norm.cdf == special.ndtr)."""
    u = _sp.ndtr(x)
    return sigma * np.sqrt(-2 * np.log(1 - u))


def rayleigh_to_normal(x, sigma=1.0):
    """utils.py:416-421 (scipy.stats.norm.ppf == special.ndtri)."""
    u = 1 - np.exp(-(x ** 2) / (2 * sigma ** 2))
    u = min(max(1e-6, u), 1 - 1e-6)
    return _sp.ndtri(u)


# ------------------------------------------------------------------ Environment
def polygon_contains_strict(verts, p):
    """shapely ``Polygon(verts).contains(Point(p))`` (Environment.py:810-817) restated as an even-odd ray cast;
    points on an edge or a vertex are not inside.  shapely / GEOS is a third-party dependency absent from
    /root/reference and from this image, so this predicate is pinned only by oracle/ref_shim.py's identical stub
    ("parity unpinned" for points within rounding of an edge; the path never evaluates such points in the fixtures)."""
    x, y = float(p[0]), float(p[1])
    n = len(verts)
    inside = False
    for i in range(n):
        x0, y0 = float(verts[i][0]), float(verts[i][1])
        x1, y1 = float(verts[(i + 1) % n][0]), float(verts[(i + 1) % n][1])
        cross = (x1 - x0) * (y - y0) - (y1 - y0) * (x - x0)
        if cross == 0.0 and min(x0, x1) <= x <= max(x0, x1) and min(y0, y1) <= y <= max(y0, y1):
            return False
        if (y0 > y) != (y1 > y):
            xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
            if x < xi:
                inside = not inside
    return inside


class OracleEnvironment:
    """The subset of Environment.__init__ (Environment.py:77-209) the path needs: a 2D environment, either the
    rectangular box -- solid (boundary walls first, then user walls; Environment.py:137-144, add_wall :330-342) or
    periodic (no boundary walls: only the `solid` branch builds them, :130-144) -- or a polygon `boundary` with
    optional `holes` (solid; hole walls after the `walls` param, :146-160)."""

    def __init__(self, scale=1.0, aspect=1.0, walls=(), boundary_conditions="solid", boundary=None, holes=()):
        self.is_rectangular = boundary is None
        b = [[0, 0], [aspect * scale, 0], [aspect * scale, scale], [0, scale]] if boundary is None else [list(v) for v in boundary]
        self.boundary = b
        self.holes = [[list(v) for v in h] for h in holes]
        self.boundary_conditions = boundary_conditions
        if boundary_conditions == "solid":
            self.walls = np.array([[b[(i + 1) if (i + 1) < len(b) else 0], b[i]] for i in range(len(b))], dtype=float)
        else:
            assert self.is_rectangular and not self.holes
            self.walls = np.zeros((0, 2, 2))
        for w in walls:
            self.add_wall(w)
        for h in self.holes:
            for i in range(len(h)):
                self.add_wall([h[(i + 1) if (i + 1) < len(h) else 0], h[i]])
        xs, ys = [c[0] for c in b], [c[1] for c in b]
        self.extent = np.array([min(xs), max(xs), min(ys), max(ys)], dtype=float)
        self.scale = scale
        self.aspect = aspect

    def add_wall(self, wall):
        self.walls = np.concatenate((self.walls, np.asarray(wall, dtype=float).reshape(1, 2, 2)), axis=0)

    def contains(self, pos):
        """Environment.py:781-818: four strict compares for a rectangle without holes (:793-806), else strictly
        inside the boundary polygon and strictly inside no hole (:807-817)."""
        if self.is_rectangular and not self.holes:
            e = self.extent
            return bool((pos[0] > e[0]) and (pos[0] < e[1]) and (pos[1] > e[2]) and (pos[1] < e[3]))
        is_in = polygon_contains_strict(self.boundary, pos)
        for h in self.holes:
            is_in = is_in and not polygon_contains_strict(h, pos)
        return bool(is_in)

    def apply_boundary_conditions(self, pos):
        """Environment.py:855-894, rectangular branch: solid clamp (:880-889) / periodic wrap (:877-879).
        Holes / polygon boundary: the reference re-draws a random position (:890-893) -- "in theory, this isn't
        used"; the fixtures never reach it and the oracle refuses to guess the RNG tape."""
        if self.contains(pos):
            return pos
        if not (self.is_rectangular and not self.holes):
            raise RuntimeError("position left a polygon / holed environment: the reference would re-draw it at random")
        e = self.extent
        pos = np.array(pos, dtype=float)
        if self.boundary_conditions == "periodic":
            pos[0] = pos[0] % e[1]
            pos[1] = pos[1] % e[3]
        else:
            pos[0] = min(max(pos[0], e[0] + 0.01), e[1] - 0.01)
            pos[1] = min(max(pos[1], e[2] + 0.01), e[3] - 0.01)
        return pos

    def vectors_between(self, pos1, pos2):
        """Environment.get_vectors_between___accounting_for_environment (Environment.py:657-675):
        pairwise pos1 - pos2 (utils.py:213), wrapped through the boundary when periodic (note: the
        reference compares against ``scale`` on both axes)."""
        pos1 = np.asarray(pos1, dtype=float).reshape(-1, 2)
        pos2 = np.asarray(pos2, dtype=float).reshape(-1, 2)
        v = pos1[:, None, :] - pos2[None, :, :]
        if self.boundary_conditions == "periodic":
            flip = np.abs(v) > (self.scale / 2)
            v[flip] = -np.sign(v[flip]) * (self.scale - np.abs(v[flip]))
        return v


# ------------------------------------------------------------------------ Agent
AGENT_DEFAULTS = dict(                       # Agent.py:68-84
    dt=0.05, speed_coherence_time=0.7, speed_mean=0.08, speed_std=0.08,
    rotational_velocity_coherence_time=0.08, rotational_velocity_std=120 * (np.pi / 180),
    head_direction_smoothing_timescale=0.15, thigmotaxis=0.5,
    wall_repel_distance=0.1, wall_repel_strength=1.0,
)


class OracleAgent:
    """State container mirroring Agent.__init__ (Agent.py:100-141); positions /
    velocities are set explicitly by the caller (no sampling here)."""

    def __init__(self, env, pos, velocity, params=None):
        self.env = env
        for k, v in AGENT_DEFAULTS.items():
            setattr(self, k, v)
        for k, v in (params or {}).items():
            setattr(self, k, v)
        self.t = 0.0
        self.pos = np.array(pos, dtype=float)
        self.velocity = np.array(velocity, dtype=float)
        self.rotational_velocity = 0.0
        self.prev_pos = self.pos.copy()
        self.measured_velocity = self.velocity.copy()
        self.measured_rotational_velocity = 0.0
        self.prev_measured_velocity = self.measured_velocity.copy()
        self.head_direction = self.velocity / np.linalg.norm(self.velocity)
        self.distance_travelled = 0.0
        self.distance_to_closest_wall = np.inf
        self.history = {k: [] for k in ("t", "pos", "distance_travelled", "vel", "rot_vel", "head_direction")}
        self.save_history = True

    # -- Agent.py:160-242 (random-motion branch only) --------------------------------
    def update(self, rng, dt=None, drift_velocity=None, drift_to_random_strength_ratio=1, **kwargs):
        dt = dt or self.dt
        self.dt = dt
        self.t += dt
        self.pos = np.array(self.pos, dtype=float)
        self.velocity = np.array(self.velocity, dtype=float)
        self.prev_pos = self.pos.copy()
        self.prev_velocity = self.velocity.copy()
        self.prev_measured_velocity = self.measured_velocity.copy()
        info = {"collisions": [], "first_hit": []}

        self._stochastic_velocity_update(rng, **kwargs)
        self._drift_velocity_update(rng, drift_velocity, drift_to_random_strength_ratio)
        self._wall_velocity_update(rng, **kwargs)
        self.pos += self.velocity * dt                               # Agent.py:216
        self._check_and_handle_wall_collisions(rng, info)
        if not self.env.contains(self.pos):                          # Agent.py:221-222
            self.pos = self.env.apply_boundary_conditions(self.pos)
        self._measure_velocity_of_step_taken(rng)
        self._update_head_direction()
        self._update_distance_travelled()
        if self.save_history:
            self._save_to_history()
        return info

    def _stochastic_velocity_update(self, rng, **kwargs):
        """Agent.py:268-312 (2D)."""
        w_std = kwargs.get("rotational_velocity_std", self.rotational_velocity_std)
        w_tau = kwargs.get("rotational_velocity_coherence_time", self.rotational_velocity_coherence_time)
        w_drift = kwargs.get("rotational_velocity_drift", 0)
        v_tau = kwargs.get("speed_coherence_time", self.speed_coherence_time)
        speed_mean = kwargs.get("speed_mean", self.speed_mean)
        self.rotational_velocity += ornstein_uhlenbeck(
            self.dt, self.rotational_velocity, w_drift, w_std, w_tau, rng)
        self.velocity = rotate(self.velocity, self.rotational_velocity * self.dt)
        speed = np.linalg.norm(self.velocity)
        if speed == 0:
            self.velocity, speed = 1e-8 * np.array([1, 0]), 1e-8
        z = rayleigh_to_normal(speed, sigma=speed_mean)
        z += ornstein_uhlenbeck(self.dt, z, 0, 1, v_tau, rng)
        speed_new = normal_to_rayleigh(z, sigma=speed_mean)
        if self.speed_std == 0:
            speed_new = speed_mean
        self.velocity = (speed_new / speed) * self.velocity

    def _drift_velocity_update(self, rng, drift_velocity, ratio):
        """Agent.py:324-341: deterministic pull towards drift_velocity.  The
        reference still draws a (2,) normal (x0 noise) -- kept for the RNG tape."""
        if drift_velocity is None:
            return
        if isinstance(rng, TapeRNG):
            rng = _ZeroDraw()       # the draw is multiplied by sigma == 0: nothing to pop
        self.velocity += ornstein_uhlenbeck(self.dt, self.velocity, np.asarray(drift_velocity, dtype=float),
                                            0, self.speed_coherence_time / ratio, rng)

    def _wall_velocity_update(self, rng, **kwargs):
        """Agent.py:343-421 + Environment.vectors_from_walls (Environment.py:843-853)."""
        strength = kwargs.get("wall_repel_strength", self.wall_repel_strength)
        d = kwargs.get("wall_repel_distance", self.wall_repel_distance)
        thig = kwargs.get("thigmotaxis", self.thigmotaxis)
        walls = self.env.walls
        if strength == 0.0 or len(walls) == 0:
            return
        vecs = shortest_vectors_from_points_to_lines(self.pos, walls, rng)[0]   # (W,2)
        x = np.linalg.norm(vecs, axis=-1)
        self.distance_to_closest_wall = np.min(x)
        unit = vecs / x[:, None]
        v = strength * self.speed_mean
        k = v ** 2 / d ** 2
        near = x <= d
        acc = np.where(near, k * (d - x), 0.0)                                   # Agent.py:389-393
        self.velocity += 3 * ((1 - thig) ** 2) * ((acc[:, None] * unit).sum(axis=0) * self.dt)
        with np.errstate(invalid="ignore"):
            spd = np.where(near, v * (1 - np.sqrt(1 - (d - x) ** 2 / d ** 2)), 0.0)  # Agent.py:406-409
        self.pos += 6 * (thig ** 2) * ((spd[:, None] * unit).sum(axis=0) * self.dt)

    def _check_and_handle_wall_collisions(self, rng, info):
        """Agent.py:423-441 + Environment.check_wall_collisions (Environment.py:820-841)."""
        walls = self.env.walls
        if len(walls) == 0:                                   # Environment.py:833-835: nothing to collide with
            return
        while True:
            step = np.array([self.prev_pos, self.pos])
            hit = vector_intercepts(walls, step, rng, return_collisions=True).reshape(-1)
            info["collisions"].append(hit.copy())
            if True not in hit:
                return
            first = int(np.argwhere(hit == True)[0][0])                           # noqa: E712  Agent.py:437
            info["first_hit"].append(first)
            self.velocity = wall_bounce(self.velocity, walls[first])
            self.velocity = (0.5 * self.speed_mean / np.linalg.norm(self.velocity)) * self.velocity
            self.pos = self.prev_pos + self.velocity * self.dt

    def _measure_velocity_of_step_taken(self, rng):
        """Agent.py:444-472."""
        self.measured_velocity = self.env.vectors_between(self.pos, self.prev_pos).reshape(-1) / self.dt
        if np.linalg.norm(self.measured_velocity) == 0:
            self.measured_velocity = 1e-8 * rng.randn(2)
        now, before = get_angle(self.measured_velocity), get_angle(self.prev_measured_velocity)
        self.measured_rotational_velocity = float(pi_domain(now - before)) / self.dt

    def _update_head_direction(self):
        """Agent.py:474-500 (tau is always the attribute in 2D, :489)."""
        dt, tau = self.dt, self.head_direction_smoothing_timescale
        immediate = self.measured_velocity / np.linalg.norm(self.measured_velocity)
        if tau <= dt:
            self.head_direction = immediate
            return
        hd = self.head_direction * (1 - dt / tau) + dt / tau * immediate
        self.head_direction = hd / np.linalg.norm(hd)

    def _update_distance_travelled(self):
        """Agent.py:502-507."""
        self.distance_travelled += np.linalg.norm(self.env.vectors_between(self.pos, self.prev_pos), axis=-1)[0][0]

    def _save_to_history(self):
        """Agent.py:509-521."""
        h = self.history
        h["t"].append(self.t)
        h["pos"].append(self.pos.tolist())
        h["distance_travelled"].append(self.distance_travelled)
        h["vel"].append(self.measured_velocity.tolist())
        h["head_direction"].append(self.head_direction.tolist())
        h["rot_vel"].append(self.measured_rotational_velocity)


# ---------------------------------------------------------------------- Neurons
def distances_accounting_for_environment(env, pos1, pos2, wall_geometry, rng):
    """Environment.py:677-779 for a solid 2D box -> (N1,N2) distances."""
    pos1 = np.asarray(pos1, dtype=float).reshape(-1, 2)
    pos2 = np.asarray(pos2, dtype=float).reshape(-1, 2)
    vec = env.vectors_between(pos1, pos2)                          # utils.py:213 (+ periodic wrap, Environment.py:670-675)
    dist = np.linalg.norm(vec, axis=-1)
    if wall_geometry == "euclidean":
        return dist
    segs = np.stack((np.broadcast_to(pos1[:, None, :], vec.shape),
                     np.broadcast_to(pos2[None, :, :], vec.shape)), axis=-2)   # utils.py:187-200
    if wall_geometry == "line_of_sight":
        inner = env.walls[4:]                                      # Environment.py:715-717
        blocked = vector_intercepts(segs.reshape(-1, 2, 2), inner, rng, return_collisions=True)
        blocked = (blocked.sum(axis=-1) != 0).reshape(dist.shape)
        dist = dist.copy()
        dist[blocked] = 1000                                       # Environment.py:730
        return dist
    if wall_geometry == "geodesic":
        assert len(env.walls) <= 5                                 # Environment.py:736-739
        if len(env.walls) == 4:
            return dist
        wall = env.walls[4]
        via = []
        for end in wall:                                           # Environment.py:746-752
            if env.contains(end):
                e = end.reshape(1, 2)
                via.append(np.linalg.norm(pos1[:, None, :] - e[None, :, :], axis=-1)
                           + np.linalg.norm(e[:, None, :] - pos2[None, :, :], axis=-1))
        via = np.array(via)
        blocked = vector_intercepts(segs.reshape(-1, 2, 2), wall[None], rng,
                                    return_collisions=True).reshape(dist.shape)
        dist = dist.copy()
        dist[blocked] = np.amin(via, axis=0)[blocked]              # Environment.py:769-773
        return dist
    raise ValueError(wall_geometry)


def place_cells_get_state(env, centres, widths, pos, rng, description="gaussian",
                          wall_geometry="euclidean", min_fr=0.0, max_fr=1.0, scalar_width=None):
    """PlaceCells.get_state, Neurons.py:936-981 -> (N, n_pos)."""
    dist = distances_accounting_for_environment(env, centres, pos, wall_geometry, rng)
    w = np.asarray(widths, dtype=float)[:, None]
    if description == "gaussian":
        fr = np.exp(-(dist ** 2) / (2 * (w ** 2)))
    elif description == "gaussian_threshold":
        fr = np.maximum(np.exp(-(dist ** 2) / (2 * (w ** 2))) - np.exp(-1 / 2), 0) / (1 - np.exp(-1 / 2))
    elif description == "diff_of_gaussians":
        ratio = 1.5
        fr = np.exp(-(dist ** 2) / (2 * (w ** 2))) - (1 / ratio ** 2) * np.exp(
            -(dist ** 2) / (2 * ((ratio * w) ** 2)))
        fr *= ratio ** 2 / (ratio ** 2 - 1)
    elif description == "one_hot":
        closest = np.argmin(np.abs(dist), axis=0)
        fr = np.eye(len(w))[closest].T
    elif description == "top_hat":
        # Neurons.py:975-976 compares against the *scalar* ``self.widths`` param
        fr = 1 * (dist < (scalar_width if scalar_width is not None else w))
    else:
        raise ValueError(description)
    return fr * (max_fr - min_fr) + min_fr                          # Neurons.py:978-980


def grid_cells_w(orientations):
    """GridCells.__init__ wave directions, Neurons.py:1154-1161 -> (N,3,2)."""
    w = []
    for th in orientations:
        w1 = rotate(np.array([1, 0]), th)
        w.append(np.array([w1, rotate(w1, np.pi / 3), rotate(w1, 2 * np.pi / 3)]))
    return np.array(w)


def grid_cells_get_state(gridscales, phase_offsets, w, pos, description="rectified_cosines",
                         width_ratio=4 / (3 * np.sqrt(3)), min_fr=0.0, max_fr=1.0):
    """GridCells.get_state (2D), Neurons.py:1172-1236 -> (N, n_pos)."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 2)
    gridscales = np.asarray(gridscales, dtype=float)
    origin = gridscales.reshape(-1, 1) * np.asarray(phase_offsets, dtype=float) / (2 * np.pi)
    vecs = origin[:, None, :] - pos[None, :, :]                     # utils.py:213 (pos1 - pos2)
    k = ((2 * np.pi) / gridscales)[:, None]
    phi = [k * (vecs[..., 0] * w[:, j, 0][:, None] + vecs[..., 1] * w[:, j, 1][:, None]) for j in range(3)]
    if description == "rectified_cosines":
        fr = (1 / 3) * (np.cos(phi[0]) + np.cos(phi[1]) + np.cos(phi[2]))
        full = (1 / 3) * (2 * np.cos(np.sqrt(3) * np.pi * width_ratio / 2) + 1)   # Neurons.py:1211
        fr = fr - full
        fr = fr / (1 - full)
        fr[fr < 0] = 0
    elif description == "shifted_cosines":
        fr = (2 / 3) * ((1 / 3) * (np.cos(phi[0]) + np.cos(phi[1]) + np.cos(phi[2])) + (1 / 2))
    else:
        raise ValueError(description)
    return fr * (max_fr - min_fr) + min_fr


def bvc_test_angles(dtheta=2):
    """BoundaryVectorCells.__init__, Neurons.py:1584-1596.  Quirk kept: angle 0
    appears twice and the last angle (360-dtheta) is missing."""
    n = int(360 / dtheta)
    base = np.array([1, 0])
    dirs, angs = [base], [0]
    for i in range(n - 1):
        dirs.append(rotate(base, 2 * np.pi * i * dtheta / 360))
        angs.append(2 * np.pi * i * dtheta / 360)
    return np.array(dirs, dtype=float), np.array(angs, dtype=float)


def von_mises_peak1(theta, mu, sigma):
    """utils.von_mises(..., norm=1), utils.py:441-457."""
    kappa = 1 / (sigma ** 2)
    return np.exp(kappa * np.cos(theta - mu)) * (1 / np.exp(kappa))


def bvc_cell_fr_norm(test_angles, sigma_angles):
    """Neurons.py:1599-1604."""
    return von_mises_peak1(test_angles.reshape(1, -1), 0, np.asarray(sigma_angles, dtype=float).reshape(-1, 1)).sum(axis=1)


def boundary_vector_preference(x):
    """Neurons.py:1746-1778 (np.piecewise on a (...,2) array with (...)-shaped
    conditions; later conditions overwrite earlier ones; untouched entries 0)."""
    la, lb = x[..., 0], x[..., 1]
    pref = np.zeros(la.shape)
    pos = la > 0
    with np.errstate(divide="ignore"):
        pref[pos] = 1 / la[pos]
    pref[la < 0] = -1
    pref[lb < 0] = -1
    pref[lb > 1] = -1
    return pref


def bvc_first_wall_distances(env, pos, test_directions, rng):
    """Neurons.py:1651-1684 -> (dist_to_first_wall (n_pos,T), wall id (n_pos,T))."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 2)
    n_pos, T = pos.shape[0], test_directions.shape[0]
    segs = np.tile(pos[:, None, None, :], (1, T, 2, 1))
    segs[:, :, 1, :] += test_directions[None, :, :]
    ic = vector_intercepts(segs.reshape(-1, 2, 2), env.walls, rng).reshape(n_pos, T, len(env.walls), 2)
    first = np.argmax(boundary_vector_preference(ic), axis=-1)
    d = np.take_along_axis(ic[..., 0], first[..., None], axis=-1)[..., 0]
    return d, first


def bvc_get_state(env, tuning_distances, tuning_angles, sigma_distances, sigma_angles, pos, rng,
                  dtheta=2, min_fr=0.0, max_fr=1.0, return_aux=False, head_direction=None):
    """BoundaryVectorCells.get_state, Neurons.py:1617-1744 -> (N, n_pos).  ``head_direction`` None =
    allocentric; a 2-vector = egocentric frame (Neurons.py:1693-1708: the test angles are shifted by
    utils.get_angle(head_direction), one head direction for all positions of the call)."""
    dirs, angs = bvc_test_angles(dtheta)
    norm = bvc_cell_fr_norm(angs, sigma_angles)
    d, first = bvc_first_wall_distances(env, pos, dirs, rng)        # (n_pos,T)
    mu_d = np.asarray(tuning_distances, dtype=float)[:, None, None]
    sg_d = np.asarray(sigma_distances, dtype=float)[:, None, None]
    mu_t = np.asarray(tuning_angles, dtype=float)[:, None, None]
    sg_t = np.asarray(sigma_angles, dtype=float)[:, None, None]
    g = np.exp(-((d[None] - mu_d) ** 2) / (2 * sg_d ** 2))          # utils.gaussian(norm=1), utils.py:424-438
    test_angles = np.tile(angs[None, None, :], (len(norm), d.shape[0], 1))
    if head_direction is not None:
        test_angles -= get_angle(head_direction)                    # Neurons.py:1707-1708
    vm = von_mises_peak1(test_angles, mu_t, sg_t)
    fr = (g * vm).sum(axis=-1) / norm[:, None]
    fr = fr * (max_fr - min_fr) + min_fr
    if return_aux:
        return fr, d, first
    return fr


def ovc_get_state(env, objects, object_types, tuning_distances, tuning_angles, sigma_distances, sigma_angles,
                  tuning_types, pos, rng, wall_geometry="line_of_sight", head_direction=None, min_fr=0.0, max_fr=1.0):
    """ObjectVectorCells.get_state, Neurons.py:1989-2113 -> (N_cells, N_pos).  `head_direction` (2,) or (N_pos,2)
    makes the cells egocentric (Neurons.py:2030-2047); angles in radians."""
    objects = np.asarray(objects, dtype=float).reshape(-1, 2)
    pos = np.asarray(pos, dtype=float).reshape(-1, 2)
    n_pos, n_cells, n_obj = len(pos), len(tuning_distances), len(objects)
    if n_obj == 0:
        return np.zeros((n_cells, n_pos))
    dist = distances_accounting_for_environment(env, pos, objects, wall_geometry, rng)           # (N_pos, N_obj)
    vec = env.vectors_between(pos, objects)                                                      # pos1 - pos2
    flat = -1 * vec.reshape(-1, 2)
    bearings = np.mod(np.arctan2(flat[:, 1], flat[:, 0] + 1e-6), 2 * np.pi).reshape(n_pos, n_obj)   # utils.get_angle, is_array
    if head_direction is not None:
        hd = np.asarray(head_direction, dtype=float)
        if hd.ndim == 1:
            bearings = bearings - get_angle(hd)
        else:                                                       # one head direction per position (batched agents)
            bearings = bearings - np.mod(np.arctan2(hd[:, 1], hd[:, 0] + 1e-6), 2 * np.pi)[:, None]
    d = dist[:, :, None]
    b = bearings[:, :, None]
    td, ta = np.asarray(tuning_distances, dtype=float)[None, None, :], np.asarray(tuning_angles, dtype=float)[None, None, :]
    sd, sa = np.asarray(sigma_distances, dtype=float)[None, None, :], np.asarray(sigma_angles, dtype=float)[None, None, :]
    g = np.exp(-((d - td) ** 2) / (2 * sd ** 2)) * 1                 # utils.gaussian(..., norm=1), utils.py:424-438
    fr = g * von_mises_peak1(b, ta, sa)                              # utils.von_mises(..., norm=1), utils.py:441-457
    mask = (np.asarray(object_types)[:, None] == np.asarray(tuning_types)[None, :]).astype(int)[None, :, :]
    fr = (fr * mask).sum(axis=1).T
    return fr * (max_fr - min_fr) + min_fr


def bin_data_for_histogramming(data, extent, dx, weights=None, norm_by_bincount=False, return_zero_bins=False):
    """utils.bin_data_for_histogramming, 2D branch (utils.py:574-589)."""
    data = np.asarray(data, dtype=float)
    bins_x = np.arange(extent[0], extent[1] + dx, dx)
    bins_y = np.arange(extent[2], extent[3] + dx, dx)
    heatmap, xedges, yedges = np.histogram2d(data[:, 0], data[:, 1], bins=[bins_x, bins_y], weights=weights)
    zero_bins = None
    if norm_by_bincount:
        bincount, xedges, yedges = np.histogram2d(data[:, 0], data[:, 1], bins=[bins_x, bins_y])
        zero_bins = (bincount == 0)
        bincount[zero_bins] = 1
        heatmap = heatmap / bincount
    heatmap = heatmap.T[::-1, :]
    if return_zero_bins:
        return (heatmap, zero_bins.T[::-1, :])
    return heatmap


def diverging_radial_assembly(distance_range=(0.01, 0.2), angle_range=(0, 90), spatial_resolution=0.04, beta=5):
    """utils.create_diverging_radial_assembly, utils.py:1073-1112 -> (mu_d, mu_theta, sigma_d, sigma_theta)."""
    fov = [a * np.pi / 180 for a in angle_range]
    mu_d, mu_t, sg_d, sg_t = [], [], [], []
    radius = max(0.01, distance_range[0])
    xi = spatial_resolution - radius / beta
    while radius < distance_range[1]:
        res = xi + radius / beta
        dth = res / radius
        if dth / 2 > fov[1]:
            right = np.array([fov[0] + dth / 2])
        else:
            right = np.arange(fov[0] + dth / 2, fov[1], dth)
        thetas = np.concatenate((-right[::-1], right))
        for th in thetas:
            mu_d.append(radius); mu_t.append(th); sg_d.append(res); sg_t.append(res / radius)
        radius = (2 * radius + res + xi) / (2 - 1 / beta)
    return np.array(mu_d), np.array(mu_t), np.array(sg_d), np.array(sg_t)


class OracleNeurons:
    """Neurons.update / save_to_history (Neurons.py:145-171, :681-687) around a
    ``get_state(pos) -> (N, n_pos)`` callable."""

    def __init__(self, agent, n, get_state, noise_std=0.0, noise_coherence_time=0.5, save_history=True):
        self.agent, self.n, self._get_state = agent, n, get_state
        self.noise_std, self.noise_coherence_time = noise_std, noise_coherence_time
        self.firingrate = np.zeros(n)
        self.noise = np.zeros(n)
        self.save_history = save_history
        self.history = {"t": [], "firingrate": [], "spikes": []}

    def update(self, rng):
        self.noise = self.noise + ornstein_uhlenbeck(self.agent.dt, self.noise, 0, self.noise_std,
                                                     self.noise_coherence_time, _NoiseView(rng))
        if np.isnan(self.agent.pos[0]):
            fr = np.zeros(self.n)
        else:
            fr = self._get_state(self.agent.pos, rng)
        self.firingrate = fr.reshape(-1) + self.noise
        if self.save_history:
            spikes = rng.uniform((self.n,)) < (self.agent.dt * self.firingrate)
            self.history["t"].append(self.agent.t)
            self.history["firingrate"].append(self.firingrate.tolist())
            self.history["spikes"].append(spikes.tolist())


class _ZeroDraw:
    def ou_normal(self, dt, shape):
        return np.zeros(shape)


class _NoiseView:
    """Routes the (N,) OU draw of Neurons.update to the noise tape of a TapeRNG."""

    def __init__(self, rng):
        self.rng = rng

    def ou_normal(self, dt, shape):
        if isinstance(self.rng, TapeRNG):
            if self.rng.noise_xi is None:
                return np.zeros(shape)
            xi = np.asarray(self.rng.noise_xi[self.rng._noise_i], dtype=float).reshape(shape)
            self.rng._noise_i += 1
            return dt * xi
        return self.rng.ou_normal(dt, shape)
