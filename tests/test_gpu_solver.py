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


def test_drqn_solve_runs_and_is_not_worse_than_the_reference_bar(lib):
    """test/runtests.jl:114-129 "TestMDP DRQN": Chain(flattenbatch, LSTM(25, 8), Dense(8, 4)), recurrence = true, double-Q; the reference
    asserts only r >= 0.  Exercises the EpisodeReplayBuffer host mirror, the recurrent batch_train! and acting with a carried hidden state."""
    mdp = TestMDP((5, 5), 1, 6)
    rng = np.random.default_rng(1)
    model = lib.Chain(lib.flattenbatch(), lib.LSTM(25, 8, rng=rng), lib.Dense(8, len(mdp.actions()), rng=rng))
    max_steps = 6000
    exploration = lib.EpsGreedyPolicy(mdp, lib.LinearDecaySchedule(start=1.0, stop=0.01, steps=max_steps / 2), rng=rng)
    solver = lib.DeepQLearningSolver(qnetwork=model, max_steps=max_steps, learning_rate=0.005, exploration_policy=exploration,
                                     eval_freq=2000, num_ep_eval=50, log_freq=500, double_q=True, dueling=False, recurrence=True,
                                     rng=rng, verbose=False, seed=3)
    policy = lib.solve(solver, mdp)
    r_drqn = evaluate(mdp, policy, lib.action)
    assert r_drqn >= 0.0, r_drqn
    policy.engine.close()


def test_batched_evaluation_equals_basic_evaluation(lib):
    """src/evaluation_policy.jl:17-42 over many environments at once (one dqn_act call per step for all running episodes)."""
    mdp = TestMDP((5, 5), 4, 6)
    rng = np.random.default_rng(2)
    model = lib.Chain(lib.flattenbatch(), lib.Dense(100, 8, lib.tanh, rng=rng), lib.Dense(8, 4, rng=rng))
    exploration = lib.EpsGreedyPolicy(mdp, 0.5, rng=rng)
    solver = lib.DeepQLearningSolver(qnetwork=model, max_steps=600, learning_rate=0.005, exploration_policy=exploration, rng=rng, verbose=False)
    policy = lib.solve(solver, mdp)
    a = lib.basic_evaluation(policy, mdp, 7, 100, False)
    b = lib.batched_evaluation(policy, mdp, 7, 100, False)
    assert abs(a[0] - b[0]) < 1e-9 and a[1] == b[1]
    policy.engine.close()
