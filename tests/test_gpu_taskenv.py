"""The batched SpatialGoalEnvironment.step against the LIVE reference's
contribs/TaskEnvironment.py -- golden tests/golden/taskenv.npz written by oracle/gen_taskenv_golden.py -- and the
reference's own acceptance test (tests/test_taskenv.py:88-126: no NaN, goals reached under goal-vector drift).  GPU only."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "taskenv.npz")


def test_rewards_and_terminations_match_the_live_reference():
    """Fed the reference agents' positions step by step, the device-side rules (reward decay / expiry, line-of-sight goal
    test through the top_hat place-cell kernel, goal removal, reward attachment) reproduce the reference's rewards to 1e-12
    and its terminated flags exactly, for 8 independent single-agent tasks with two goals each and an inner wall."""
    from ratinabox_b200.contribs import SpatialGoalEnvironment
    g = np.load(GOLD)
    T, E = g["reward"].shape
    np.random.seed(0)
    env = SpatialGoalEnvironment(n_agents=E, dt=float(g["dt"]), possible_goal_positions=g["goals"], reset_n_goals=2,
                                 goal_radius=float(g["radius"]))
    for w in g["walls"]:
        env.add_wall(w)
    env._goal_cells._sig = None                         # (walls were added after construction: re-pack on next use)
    env.reset(goal_indices=g["goal_idx"])
    n_reached = 0
    for t in range(T):
        reached = env._apply_rules(float(g["dt"]), positions=g["pos"][t])
        n_reached += int(reached.sum())
        reward = (env.reward_state * env.reward_on).sum(dim=1).cpu().numpy()
        terminated = (~env.goal_active.any(dim=1)).cpu().numpy()
        assert np.abs(reward - g["reward"][t]).max() <= 1e-12, (t, reward, g["reward"][t])
        assert np.array_equal(terminated, g["terminated"][t]), t
    assert n_reached == 2 * E and g["terminated"][-1].all()


@pytest.mark.parametrize("A,n_goals", [(1, 1), (3, 2), (1000, 2)])
def test_agents_reach_their_goals_under_goal_vector_drift(A, n_goals):
    """The reference's acceptance test for its RL wrapper, batched: drift = goal vector x speed_mean x const
    (tests/test_taskenv.py:104-108); positions, actions and rewards never NaN; every agent finishes within 10 000 steps."""
    from ratinabox_b200.contribs import SpatialGoalEnvironment
    np.random.seed(3)
    env = SpatialGoalEnvironment(n_agents=A, dt=0.01, possible_goal_positions="random_5", reset_n_goals=n_goals)
    env.reset()                                          # (an open box like the reference's fixtures: the straight-line policy
                                                         # of its test cannot route around walls; the golden test has one)
    done = np.zeros(A, dtype=bool)
    total = np.zeros(A)
    for step in range(10_000):
        action = env.get_goal_vectors() * env.Ag.speed_mean * 8.0
        assert not np.isnan(action).any()
        obs, reward, terminated, truncated, info = env.step(action)
        assert obs.shape == (A, 2) and not np.isnan(obs).any() and not np.isnan(reward).any()
        assert reward.min() >= 0.0 and reward.max() <= n_goals * 1.0 + 1e-12
        total += reward
        done = terminated
        if done.all():
            break
    assert done.all(), (step, done.mean())
    assert (total > 0).all()                             # every agent collected reward on its way
