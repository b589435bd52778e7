"""Batched spatial-goal task: the policy-control caller of the hot path.

The reference's ``ratinabox.contribs.TaskEnvironment.SpatialGoalEnvironment`` (a pettingzoo ``ParallelEnv``) steps its
agents in a Python loop -- ``Ag.update(dt=dt, drift_velocity=action, drift_to_random_strength_ratio=strength)`` per
agent (contribs/TaskEnvironment.py:399-408) -- then decays the active rewards, checks the goals and returns
``(observations, rewards, terminated, truncated, infos)`` (:410-447).  This class is that ``step()`` for a BATCH of
``n_agents`` independent single-agent tasks that share one Environment geometry and one pool of goal positions:

* ``step(actions)``: actions ``(n_agents, 2)`` on the host (NaN -> 0, contribs/TaskEnvironment.py:402-404) are the
  agents' ``drift_velocity``; one batched ``Agent.update`` (the CUDA motion kernel) + the attached observation
  populations' rate kernels; then, on the device,
* rewards decay like ``Reward.update`` (:804-823): ``state += -decay(state) * dt`` (preset "linear": ``decay = knob * state``,
  "constant": ``knob``, "none": 0), ``expire_clock -= dt``, a reward is dropped when its clock reaches 0;
* a goal is reached when the agent is within ``goal_radius`` of it under the reference's ``line_of_sight`` metric
  (``SpatialGoal._in_goal_radius``, :1319-1332: ``Environment.get_distances_between___accounting_for_environment``) --
  evaluated by the engine's own ``top_hat`` PlaceCells kernel with the goals as centres, so walls block goals exactly as
  they block place fields; a reached goal is removed from that agent's list and a fresh copy of the goal's reward
  (``Reward(1, dt, expire_clock=1, decay="linear")``, :948) is attached to the agent (:283-285);
* ``reward`` = sum of the agent's active reward states (``RewardCache.get_total``, :929-936); ``terminated`` = the agent has
  no goals left (:289; every task here has ONE agent, for which the reference's "interact" and "noninteract" agent modes
  coincide).

Not reproduced (they are Python-object bookkeeping of the pettingzoo wrapper, not arithmetic of the step): the
agents/observation/action ``spaces`` registries, rendering, episode statistics, sequential goal order, ``TimeElapsedGoal``,
external reward drives, and the reference's remove-while-iterating quirk in ``RewardCache.update`` (:913-925) that lets the
second of two rewards expiring in the same step live one step longer.  Terminated agents keep moving with the batch; call
``reset(mask=terminated)`` to give them new goals.  pettingzoo / gymnasium are not needed (nor installed here)."""
import numpy as np

from ..Environment import Environment
from ..Agent import Agent
from ..Neurons import PlaceCells


class SpatialGoalEnvironment(Environment):
    DECAYS = ("linear", "constant", "none")

    def __init__(self, params={}, n_agents=1, dt=0.01, possible_goal_positions="random_5", reset_n_goals=1,
                 goal_radius=None, reward_value=1.0, reward_expire_clock=1.0, reward_decay="linear", reward_decay_knob=None,
                 teleport_on_reset=False, agent_params={}):
        super().__init__(params)
        if reward_decay not in self.DECAYS:
            raise ValueError(f"reward_decay must be one of {self.DECAYS}")
        self.dt = dt
        self.n_agents = int(n_agents)
        self.teleport_on_reset = teleport_on_reset
        # contribs/TaskEnvironment.py:1417-1456: a list / array of positions, or "random_<n>"
        if isinstance(possible_goal_positions, str):
            assert possible_goal_positions.startswith("random"), "possible_goal_positions: positions or 'random_<n>'"
            n = int(possible_goal_positions.split("_")[1]) if "_" in possible_goal_positions else 5
            ext = np.asarray(self.extent, dtype=float)
            lo, hi = ext[[0, 2]], ext[[1, 3]]
            self.goal_positions = lo + np.random.rand(n, 2) * (hi - lo)
        else:
            self.goal_positions = np.asarray(possible_goal_positions, dtype=float).reshape(-1, 2)
        self.n_goals = len(self.goal_positions)
        self.reset_n_goals = int(reset_n_goals)
        if not (0 < self.reset_n_goals <= self.n_goals):
            raise ValueError("reset_n_goals must be in [1, number of possible goals]")
        # SpatialGoal.__init__ (:1313-1317)
        self.goal_radius = float(np.min((self.dx * 10, np.ptp(self.extent) / 10)) if goal_radius is None else goal_radius)
        self.reward_value, self.reward_expire_clock = float(reward_value), float(reward_expire_clock)
        self.reward_decay = reward_decay
        self.reward_decay_knob = float({"linear": 1, "constant": 1, "none": 0}[reward_decay] if reward_decay_knob is None
                                       else reward_decay_knob)                      # Reward.decay_knobs_preset (:738-743)
        self.Ag = Agent(self, dict({"dt": dt, "n_agents": self.n_agents}, **agent_params))
        # goal test = the engine's top_hat place-cell kernel over the goal positions (line_of_sight distances)
        self._goal_cells = PlaceCells(self.Ag, {"place_cell_centres": self.goal_positions, "widths": self.goal_radius,
                                                "description": "top_hat", "wall_geometry": "line_of_sight",
                                                "name": "goal_test"})
        self.Ag.Neurons.remove(self._goal_cells)           # evaluated by step(), not part of the agents' observation stack
        import torch
        self._torch = torch
        dev = self.Ag.device
        A, G = self.n_agents, self.n_goals
        self.goal_active = torch.zeros((A, G), dtype=torch.bool, device=dev)
        self.reward_state = torch.zeros((A, G), dtype=torch.float64, device=dev)
        self.reward_clock = torch.zeros((A, G), dtype=torch.float64, device=dev)
        self.reward_on = torch.zeros((A, G), dtype=torch.bool, device=dev)
        self.t = 0.0
        self.episode = 0
        self.reset()

    # ------------------------------------------------------------------ pettingzoo-shaped API
    def reset(self, seed=None, goal_indices=None, mask=None):
        """New goals (``reset_n_goals`` drawn per agent from the pool, or the given ``(n_agents, k)`` indices) for every
        agent (or those in ``mask``); rewards cleared; optionally teleport (contribs/TaskEnvironment.py:302-351)."""
        torch = self._torch
        if seed is not None:
            np.random.seed(seed)
        A, G = self.n_agents, self.n_goals
        which = np.ones(A, dtype=bool) if mask is None else np.asarray(mask, dtype=bool).reshape(A)
        if goal_indices is None:
            goal_indices = np.argsort(np.random.rand(A, G), axis=1)[:, : self.reset_n_goals]
        goal_indices = np.asarray(goal_indices).reshape(A, -1)
        active = np.zeros((A, G), dtype=bool)
        np.put_along_axis(active, goal_indices, True, axis=1)
        w = torch.as_tensor(which, device=self.Ag.device)[:, None]
        self.goal_active = torch.where(w, torch.as_tensor(active, device=self.Ag.device), self.goal_active)
        self.reward_on = self.reward_on & ~w
        self.reward_state = torch.where(w, torch.zeros_like(self.reward_state), self.reward_state)
        if self.teleport_on_reset:
            pos = np.array(self.Ag.pos, dtype=float).reshape(A, 2)
            pos[which] = self.sample_positions(n=int(which.sum()), method="random")
            self.Ag.pos = pos
        self.episode += 1
        return self.get_observation(), {}

    def get_observation(self):
        """Agent positions ``(n_agents, 2)`` (the reference's default observation, contribs/TaskEnvironment.py:205-216);
        firing rates of populations attached to ``self.Ag`` are read from them directly (``Ns.firingrate``)."""
        return np.array(self.Ag.pos, dtype=float).reshape(self.n_agents, 2)

    def get_goal_vectors(self):
        """Vector from every agent to its nearest ACTIVE goal, zeros when none is left (the reference's test helper
        ``get_goal_vector``, contribs/TaskEnvironment.py:1553-1584)."""
        pos = self.get_observation()
        act = self.goal_active.cpu().numpy()
        vec = self.goal_positions[None, :, :] - pos[:, None, :]
        d = np.linalg.norm(vec, axis=2)
        d[~act] = np.inf
        j = np.argmin(d, axis=1)
        out = vec[np.arange(self.n_agents), j]
        out[~act.any(axis=1)] = 0.0
        return out

    def step(self, actions=None, dt=None, drift_to_random_strength_ratio=1):
        """One tick of every task: see the module docstring.  Returns host arrays
        ``(observation (A,2), reward (A,), terminated (A,), truncated (A,), info)``."""
        torch = self._torch
        A = self.n_agents
        dt = self.dt if dt is None else dt
        drift = None
        if actions is not None:
            drift = np.array(actions, dtype=np.float64).reshape(A, 2)
            drift[np.isnan(drift)] = 0                                        # contribs/TaskEnvironment.py:402-404
        self.Ag.update(dt=dt, drift_velocity=drift, drift_to_random_strength_ratio=drift_to_random_strength_ratio)
        for ns in self.Ag.Neurons:
            ns.update()
        self.t += dt
        self._apply_rules(dt)
        reward = (self.reward_state * self.reward_on).sum(dim=1)
        terminated = ~self.goal_active.any(dim=1)
        return (self.get_observation(), reward.cpu().numpy(), terminated.cpu().numpy(), np.zeros(A, dtype=bool), {})

    # ------------------------------------------------------------------ rules (device tensors)
    def _apply_rules(self, dt, positions=None):
        """Reward decay, goal test, new rewards -- the arithmetic of RewardCache.update / GoalCache.check / Reward.update.
        ``positions`` (parity tap): evaluate the goal test at these positions instead of the agents' own."""
        torch = self._torch
        on = self.reward_on
        if self.reward_decay == "linear":
            delta = -(self.reward_decay_knob * self.reward_state)
        elif self.reward_decay == "constant":
            delta = torch.full_like(self.reward_state, -self.reward_decay_knob)
        else:
            delta = torch.zeros_like(self.reward_state)
        self.reward_state = torch.where(on, self.reward_state + delta * dt, self.reward_state)      # Reward.update (:818)
        self.reward_clock = torch.where(on, self.reward_clock - dt, self.reward_clock)
        self.reward_on = on & ~(self.reward_clock <= 0)
        if positions is None:
            inside = self._goal_cells.get_state(evaluate_at="agent", return_tensor=True)            # (A, G) 0 / 1
        else:
            inside = self._goal_cells.get_state(evaluate_at=None, pos=positions, return_tensor=True)
        reached = (inside > 0.5) & self.goal_active
        self.goal_active = self.goal_active & ~reached
        self.reward_state = torch.where(reached, torch.full_like(self.reward_state, self.reward_value), self.reward_state)
        self.reward_clock = torch.where(reached, torch.full_like(self.reward_clock, self.reward_expire_clock), self.reward_clock)
        self.reward_on = self.reward_on | reached
        return reached
