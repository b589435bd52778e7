"""TEST INFRASTRUCTURE -- golden vectors for the batched SpatialGoalEnvironment from the LIVE reference.

    python oracle/gen_taskenv_golden.py        # -> tests/golden/taskenv.npz   (build container: /root/reference)

Runs the unmodified ``ratinabox.contribs.TaskEnvironment.SpatialGoalEnvironment`` (pettingzoo / gymnasium are not
installed here: the two names the module needs from them -- ``pettingzoo.ParallelEnv`` as a base class and
``gymnasium.spaces.{Box, Space, Dict}`` as containers -- are stand-in classes; none of them takes part in the arithmetic
of ``step()``).  E independent single-agent tasks (one reference environment each) are driven towards their goals with
the reference's own test policy (tests/test_taskenv.py:104-108: goal vector x speed_mean x const); per step the agent's
position, the reward and the terminated flag are recorded.  The batched class is then fed the SAME positions
(``_apply_rules(dt, positions=...)``) and must reproduce rewards and terminations."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
WALLS = [[[0.5, 0.0], [0.5, 0.3]]]
GOALS = np.array([[0.15, 0.45], [0.85, 0.5], [0.8, 0.85], [0.3, 0.8]])


def install_rl_stubs():
    pz = types.ModuleType("pettingzoo")
    pz.ParallelEnv = type("ParallelEnv", (), {})
    gym, sp = types.ModuleType("gymnasium"), types.ModuleType("gymnasium.spaces")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low=None, high=None, shape=None, dtype=float):
            self.low, self.high, self.shape = low, high, shape

    class Dict(Space, dict):
        def __init__(self, d=None):
            dict.__init__(self, d or {})
            self.spaces = self

        def __class_getitem__(cls, k):
            return cls

    sp.Box, sp.Space, sp.Dict = Box, Space, Dict
    gym.spaces = sp
    sys.modules.update({"pettingzoo": pz, "gymnasium": gym, "gymnasium.spaces": sp})


def main(E=8, T=1600, speed_const=8.0):
    ref_shim.install()
    install_rl_stubs()
    assert ref_shim.import_reference() is not None, "needs /root/reference (or oracle/_ref)"
    import warnings
    warnings.simplefilter("ignore")
    from ratinabox.contribs.TaskEnvironment import SpatialGoalEnvironment, get_goal_vector
    from ratinabox.Agent import Agent
    pos = np.zeros((T, E, 2)); rew = np.zeros((T, E)); term = np.zeros((T, E), dtype=bool)
    goal_idx = np.zeros((E, 2), dtype=np.int64); start = np.zeros((E, 2)); radius = None
    for e in range(E):
        np.random.seed(100 + e)
        gi = np.random.choice(len(GOALS), 2, replace=False)
        env = SpatialGoalEnvironment(params={}, render_every=10 ** 9, teleport_on_reset=False, dt=0.01,
                                     possible_goal_positions=GOALS[gi], goalcachekws={"reset_n_goals": 2}, verbose=False)
        for w in WALLS:
            env.add_wall(np.array(w))
        ag = Agent(env, {"dt": 0.01})
        ag.pos = np.array([0.1 + 0.8 * np.random.rand(), 0.1 + 0.8 * np.random.rand()])
        env.add_agents(ag)
        env.reset()
        goal_idx[e], start[e], radius = gi, ag.pos, env.goal_cache.get_goals()[0].radius
        done = False
        for t in range(T):
            if not done:
                act = {name: v * ag.speed_mean * speed_const for name, v in get_goal_vector([ag]).items()}
                _, r, tm, _, _ = env.step(act)
                done = all(tm.values())
                rew[t, e], term[t, e] = list(r.values())[0], done
            else:
                # after termination the reference stops stepping the agent; keep decaying the rewards the way
                # step() does (reward_cache.update, contribs/TaskEnvironment.py:410-412) to pin the decay / expiry arithmetic
                for rc in env.reward_caches.values():
                    rc.update()
                rew[t, e], term[t, e] = list(env.get_reward().values())[0], True
            pos[t, e] = ag.pos
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "taskenv.npz"), walls=np.array(WALLS), goals=GOALS, goal_idx=goal_idx, start=start,
                        radius=radius, pos=pos, reward=rew, terminated=term, dt=0.01)
    print("reached:", term.any(axis=0), "steps to finish:", term.argmax(axis=0), "max reward", rew.max(), "radius", radius)


if __name__ == "__main__":
    main()
