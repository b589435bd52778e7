"""Host mirror of ``ratinabox.Environment`` for the CUDA step engine.

Only what the hot path needs lives here: a rectangular 2D box (solid
or periodic) with internal walls.  Construction semantics follow the reference:
boundary walls first, in the reference's corner order (ratinabox/Environment.py:118-144),
then user walls in insertion order (``add_wall``, :330-342).  Anything outside
that (1D, polygon boundaries, holes, objects) raises
``NotImplementedError`` instead of silently taking a different path.
"""
import copy
import warnings

import numpy as np


class Environment:
    default_params = {            # ratinabox/Environment.py:65-75
        "dimensionality": "2D",
        "boundary_conditions": "solid",
        "scale": 1,
        "aspect": 1,
        "dx": 0.01,
        "boundary": None,
        "walls": [],
        "holes": [],
        "objects": [],
    }

    def __init__(self, params={}):
        self.params = copy.deepcopy(__class__.default_params)
        self.params.update(params)
        unexpected = [k for k in params if k not in __class__.default_params]
        if unexpected:                                              # utils.check_params, utils.py:877-916
            warnings.warn(f"Found {len(unexpected)} unexpected params key(s) while initializing Environment: {unexpected}")
        for k, v in self.params.items():
            setattr(self, k, v)
        if self.dimensionality != "2D":
            raise NotImplementedError("ratinabox_b200 accelerates 2D environments only (SURVEY.md section 2 row 9)")
        if self.boundary_conditions not in ("solid", "periodic"):
            raise ValueError(f"unknown boundary_conditions {self.boundary_conditions!r}")
        if self.boundary is not None or len(self.holes) > 0:
            raise NotImplementedError("polygon boundaries / holes are outside the CUDA hot path (SURVEY.md section 2 row 11)")
        if len(self.objects) > 0:
            raise NotImplementedError("objects are outside the CUDA hot path (SURVEY.md section 2 row 13)")

        self.D = 2
        self.is_rectangular = True
        self.has_holes = False
        self.Agents = []
        self.agents_dict = {}
        b = [[0, 0], [self.aspect * self.scale, 0], [self.aspect * self.scale, self.scale], [0, self.scale]]
        self.boundary = b
        user_walls = np.array(self.walls, dtype=float).reshape(-1, 2, 2)
        if self.boundary_conditions == "solid":                     # Environment.py:137-144
            boundary_walls = np.array([[b[(i + 1) if (i + 1) < len(b) else 0], b[i]] for i in range(len(b))], dtype=float)
            self.walls = np.vstack((boundary_walls, user_walls))
            self.n_boundary_walls = 4
        else:                                                       # periodic: no boundary walls are built
            self.walls = user_walls
            self.n_boundary_walls = 0
        left, right = min(c[0] for c in b), max(c[0] for c in b)
        bottom, top = min(c[1] for c in b), max(c[1] for c in b)
        self.centre = np.array([(left + right) / 2, (top + bottom) / 2])
        self.extent = np.array([left, right, bottom, top], dtype=float)
        self.discrete_coords = self.discretise_environment(dx=self.dx)
        self.flattened_discrete_coords = self.discrete_coords.reshape(-1, self.discrete_coords.shape[-1])
        self._walls_version = 0
        self._dev = {}          # device -> (version, tensor)

    # ------------------------------------------------------------------ registry
    def add_agent(self, agent=None):                                # Environment.py:220-250
        if agent.name in self.agents_dict:
            raise ValueError(f"An agent with the name {agent.name} already exists in the environment.")
        self.Agents.append(agent)
        self.agents_dict[agent.name] = agent

    # --------------------------------------------------------------------- walls
    def add_wall(self, wall):                                       # Environment.py:330-342
        wall = np.asarray(wall, dtype=float).reshape(1, 2, 2)
        self.walls = np.concatenate((self.walls, wall), axis=0)
        self._walls_version += 1

    def _walls_signature(self):
        return (self._walls_version, self.walls.shape[0], hash(self.walls.tobytes()))

    def walls_device(self, device):
        """(W,2,2) float64 walls on ``device`` (re-uploaded when they changed)."""
        import torch
        sig = self._walls_signature()
        hit = self._dev.get(device)
        if hit is None or hit[0] != sig:
            t = torch.as_tensor(np.ascontiguousarray(self.walls, dtype=np.float64), device=device)
            self._dev[device] = (sig, t)
            return t
        return hit[1]

    # ------------------------------------------------------------------ sampling
    def sample_positions(self, n=10, method="uniform_jitter"):      # Environment.py:560-633 (2D, rectangular)
        ex = self.extent
        if method == "random":
            positions = np.zeros((n, 2))
            positions[:, 0] = np.random.uniform(ex[0], ex[1], size=n)
            positions[:, 1] = np.random.uniform(ex[2], ex[3], size=n)
            return positions
        if method[:7] == "uniform":
            area = (ex[1] - ex[0]) * (ex[3] - ex[2])
            delta = np.sqrt(area / n)
            x = np.linspace(ex[0] + delta / 2, ex[1] - delta / 2, int((ex[1] - ex[0]) / delta))
            y = np.linspace(ex[2] + delta / 2, ex[3] - delta / 2, int((ex[3] - ex[2]) / delta))
            positions = np.array(np.meshgrid(x, y)).reshape(2, -1).T
            n_uniform = positions.shape[0]
            if method[7:] == "_jitter":
                positions = positions + np.random.uniform(-0.45 * delta, 0.45 * delta, positions.shape)
            n_remaining = n - n_uniform
            if n_remaining > 0:
                extra = np.array([positions[i] for i in np.random.choice(range(len(positions)), n_remaining, replace=True)])
                delta /= 2
                extra = extra + np.random.uniform(-0.45 * delta, 0.45 * delta, extra.shape)
                positions = np.vstack((positions, extra))
            return positions
        raise ValueError(f"unknown sampling method {method!r}")

    def discretise_environment(self, dx=None):                      # Environment.py:635-655
        dx = self.dx if dx is None else dx
        minx, maxx, miny, maxy = self.extent
        self.x_array = np.arange(minx + dx / 2, maxx, dx)
        self.y_array = np.arange(miny + dx / 2, maxy, dx)[::-1]
        xm, ym = np.meshgrid(self.x_array, self.y_array)
        return np.stack((xm, ym), axis=-1)

    def check_if_position_is_in_environment(self, pos):             # Environment.py:781-818 (rectangle, no holes)
        pos = np.asarray(pos, dtype=float).reshape(-1)
        e = self.extent
        return bool((pos[0] > e[0]) and (pos[0] < e[1]) and (pos[1] > e[2]) and (pos[1] < e[3]))
