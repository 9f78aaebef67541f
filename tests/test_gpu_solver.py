"""The reference's own acceptance criterion for the path, end to end through the host mirror (test/runtests.jl:96-111, "Prioritized
DDQN"): solve(solver, TestMDP((5,5), 4, 6)) with double-Q + dueling + prioritized replay must reach an average return >= 1.5 (the
optimum is 2.1).  Every forward pass, replay write, sample, gradient step and target sync of the run goes through libdqn_b200.so."""
import numpy as np
import pytest

from test_env import TestMDP, evaluate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fp32", "3xtf32"])
def test_prioritized_ddqn_reaches_reference_threshold(lib, math_mode):
    mdp = TestMDP((5, 5), 4, 6)
    rng = np.random.default_rng(1)
    model = lib.Chain(lib.flattenbatch(), lib.Dense(100, 8, lib.tanh, rng=rng), lib.Dense(8, len(mdp.actions()), rng=rng))
    max_steps = 10000
    exploration = lib.EpsGreedyPolicy(mdp, lib.LinearDecaySchedule(start=1.0, stop=0.01, steps=max_steps / 2), rng=rng)
    solver = lib.DeepQLearningSolver(qnetwork=model, max_steps=max_steps, learning_rate=0.005, exploration_policy=exploration,
                                     eval_freq=2000, num_ep_eval=100, log_freq=500, double_q=True, dueling=True, prioritized_replay=True,
                                     rng=rng, verbose=False, math_mode=math_mode, seed=3)
    policy = lib.solve(solver, mdp)
    r_ddqn = evaluate(mdp, policy, lib.action)
    assert r_ddqn >= 1.5, r_ddqn
    assert lib.actionvalues(policy, mdp.observe()).shape == (len(mdp.actions()),)
    evals = [h["eval_reward"] for h in solver.history if "eval_reward" in h]
    assert evals and max(evals) >= 1.5
    policy.engine.close()
