"""Host mirror of the callers of the hot path: DeepQLearningSolver / solve / dqn_train! / batch_train! / NNPolicy.

Only the orchestration lives here (it stays host code in the reference too, src/solver.jl:30-189); every
numerical step is a call into libdqn_b200.so.  The environment protocol is CommonRLInterface's, spelled
in Python: reset(), actions(), observe(), act(a) -> reward, terminated(); an optional `discount` attribute
plays default_discount (src/helpers.jl:83-85).  Observations are numpy arrays whose memory image equals
the Julia array's, i.e. a Julia (W,H,C) observation is a C-ordered numpy (C,H,W)."""
import os
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import _capi
from .engine import Engine, make_config
from .flux import Chain, DuelingNetwork, create_dueling_network, flat_params, load_flat_params, isrecurrent
from .replay import PrioritizedReplayBuffer, EpisodeReplayBuffer, DQExperience


# ---- POMDPTools stand-ins used by the reference's call sites (src/solver.jl:83,155) -----------------
@dataclass
class LinearDecaySchedule:
    start: float
    stop: float
    steps: float

    def __call__(self, k):
        rate = (self.start - self.stop) / self.steps
        return max(self.stop, self.start - k * rate)


class EpsGreedyPolicy:
    def __init__(self, env, eps, rng=None):
        self.eps = eps if callable(eps) else (lambda k, e=eps: e)
        self.rng = rng or np.random.default_rng()
        self.actions = list(env.actions())

    def action(self, on_policy, k, obs):
        if self.rng.random() < self.eps(k):
            return self.actions[int(self.rng.integers(len(self.actions)))]
        return action(on_policy, obs)

    def loginfo(self, k):
        return {"eps": self.eps(k)}


# ---- src/policy.jl ---------------------------------------------------------------------------------
class NNPolicy:
    """src/policy.jl:17-22.  `qnetwork` is the host descriptor; the weights that act live in the engine."""

    def __init__(self, problem, qnetwork, action_map, n_input_dims, engine=None):
        self.problem, self.qnetwork, self.action_map, self.n_input_dims, self.engine = problem, qnetwork, list(action_map), n_input_dims, engine

    def _q(self, o):
        o = np.asarray(o)
        if o.ndim != self.n_input_dims:                       # src/policy.jl:44
            raise ValueError(f"NNPolicyError: was expecting an array with {self.n_input_dims} dimensions, got {o.ndim}")
        return self.engine.q_values(o[None])[0]


def getnetwork(policy):
    return policy.qnetwork


def resetstate(policy):
    """src/policy.jl:32-34: Flux.reset!(policy.qnetwork) - the acting hidden state of a recurrent engine goes back to state0"""
    if policy.engine is not None and policy.engine.recurrent:
        policy.engine.policy_reset()


def action(policy, o):                                         # src/policy.jl:38-46: first maximal index
    return policy.action_map[int(np.argmax(policy._q(o)))]


def actionvalues(policy, o):                                   # src/policy.jl:48-55
    return policy._q(o)


def value(policy, o):                                          # src/policy.jl:57-64
    return float(np.max(policy._q(o)))


# ---- src/evaluation_policy.jl:17-42 ----------------------------------------------------------------
def basic_evaluation(policy, env, n_eval, max_episode_length, verbose):
    avg_r = avg_steps = 0.0
    for _ in range(n_eval):
        done, r_tot, step = False, 0.0, 0
        env.reset()
        obs = env.observe()
        resetstate(policy)
        while not done and step <= max_episode_length:
            rew = env.act(action(policy, obs))
            obs = env.observe()
            done = env.terminated()
            r_tot += rew
            step += 1
        avg_steps += step
        avg_r += r_tot
    if verbose:
        print("Evaluation ... Avg Reward %2.2f | Avg Step %2.2f " % (avg_r / n_eval, avg_steps / n_eval))
    return avg_r / n_eval, avg_steps / n_eval, {}


def batched_evaluation(policy, env, n_eval, max_episode_length, verbose, make_env=None):
    """basic_evaluation (src/evaluation_policy.jl:17-42) over n_eval copies of the environment stepped in lockstep: ONE device call
    (dqn_act: forward + argmax on the GPU) picks the action of every running episode per step instead of one batch-1 forward per
    environment step.  Same episode accounting as the reference (`step <= max_episode_length`); deterministic environments give the
    same averages as basic_evaluation."""
    import copy
    envs = [make_env() if make_env else copy.deepcopy(env) for _ in range(n_eval)]
    for e in envs:
        e.reset()
    resetstate(policy)
    r_tot = np.zeros(n_eval); steps = np.zeros(n_eval, np.int64)
    running = np.array([not e.terminated() for e in envs])
    call = 0
    while running.any():
        live = np.nonzero(running)[0]
        obs = np.stack([np.asarray(envs[i].observe()) for i in live])
        acts = policy.engine.act(obs, eps=0.0, call=call)
        call += 1
        for i, ai in zip(live, acts):
            r_tot[i] += envs[i].act(policy.action_map[int(ai) - 1])
            steps[i] += 1
            if envs[i].terminated() or steps[i] > max_episode_length:
                running[i] = False
    if verbose:
        print("Evaluation ... Avg Reward %2.2f | Avg Step %2.2f " % (r_tot.mean(), steps.mean()))
    return float(r_tot.mean()), float(steps.mean()), {}


# ---- src/solver.jl:1-28 ----------------------------------------------------------------------------
@dataclass
class DeepQLearningSolver:
    exploration_policy: Any
    qnetwork: Any = None
    learning_rate: float = 1e-4
    max_steps: int = 1000
    batch_size: int = 32
    train_freq: int = 4
    eval_freq: int = 500
    target_update_freq: int = 500
    num_ep_eval: int = 100
    double_q: bool = True
    dueling: bool = True
    recurrence: bool = False
    evaluation_policy: Any = basic_evaluation
    trace_length: int = 40
    prioritized_replay: bool = True
    prioritized_replay_alpha: float = 0.6      # dead in the reference: the buffer keeps its own defaults
    prioritized_replay_epsilon: float = 1e-6   # dead in the reference
    prioritized_replay_beta: float = 0.4       # dead in the reference
    buffer_size: int = 1000
    max_episode_length: int = 100
    train_start: int = 200
    rng: Any = None
    logdir: Optional[str] = None
    save_freq: int = 3000
    log_freq: int = 100
    verbose: bool = True
    # engine knobs (no counterpart in the reference)
    device: int = 0
    obs_dtype: str = "f32"
    math_mode: int = _capi.MATH_FP32
    seed: int = 0
    history: list = field(default_factory=list)


def default_discount(env):                                     # src/helpers.jl:83-85
    return float(getattr(env, "discount", 1.0))


def populate_replay_buffer(replay, env, action_indices, max_pop=None, max_steps=100, rng=None):
    """src/prioritized_experience_replay.jl:106-134 (random policy, initial td error |r|)."""
    rng = rng or np.random.default_rng()
    acts = list(env.actions())
    max_pop = replay.max_size if max_pop is None else max_pop
    env.reset()
    o = env.observe()
    step = 0
    S, A, R, SP, D = [], [], [], [], []
    for _ in range(max_pop - replay._curr_size):
        a = acts[int(rng.integers(len(acts)))]
        rew = env.act(a)
        op = env.observe()
        done = env.terminated()
        S.append(o); A.append(action_indices[a]); R.append(np.float32(rew)); SP.append(op); D.append(done)
        o = op
        step += 1
        if done or step >= max_steps:
            env.reset()
            o = env.observe()
            step = 0
    if A:
        replay.add_batch(np.stack(S), A, R, np.stack(SP), D, np.abs(np.asarray(R, np.float32)))
    assert replay._curr_size >= replay.batch_size


def generate_episode(env, action_indices, max_steps=100, rng=None):
    """src/episode_replay.jl:108-130 (random policy)"""
    rng = rng or np.random.default_rng()
    acts = list(env.actions())
    episode = []
    env.reset()
    o = env.observe()
    done, step = False, 1
    while not done and step < max_steps:
        a = acts[int(rng.integers(len(acts)))]
        rew = env.act(a)
        op = env.observe()
        done = env.terminated()
        episode.append(DQExperience(np.asarray(o), action_indices[a], np.float32(rew), np.asarray(op), done))
        o = op
        step += 1
    return episode


def populate_episode_buffer(replay, env, action_indices, max_pop=None, max_steps=100, rng=None):
    """src/episode_replay.jl:97-106"""
    max_pop = replay.max_size if max_pop is None else max_pop
    for _ in range(max_pop - replay._curr_size):
        replay.add_episode(generate_episode(env, action_indices, max_steps=min(max_steps, replay.max_len + 1), rng=rng))
    assert replay._curr_size >= replay.batch_size


def initialize_replay_buffer(solver, env, action_indices, engine):
    """src/solver.jl:180-189."""
    if solver.recurrence:
        replay = EpisodeReplayBuffer(engine)                      # src/solver.jl:183-184
        populate_episode_buffer(replay, env, action_indices, max_pop=solver.train_start, rng=solver.rng)
        return replay
    replay = PrioritizedReplayBuffer(engine)                  # alpha, beta, eps: the constructor defaults (PER.jl:43-45)
    populate_replay_buffer(replay, env, action_indices, max_pop=solver.train_start, rng=solver.rng)
    return replay


def _chain_layers(qnetwork):
    return list(qnetwork.layers)


def solve(solver, env):
    """src/solver.jl:40-57."""
    action_map = list(env.actions())
    action_indices = {a: i + 1 for i, a in enumerate(action_map)}
    if isrecurrent(solver.qnetwork) and not solver.recurrence:
        raise ValueError("DeepQLearningError: you passed in a recurrent model but recurrence is set to false")
    env.reset()
    obs = np.asarray(env.observe())
    active_q = create_dueling_network(solver.qnetwork, rng=solver.rng) if solver.dueling else solver.qnetwork
    flux_shape = tuple(reversed(obs.shape))
    cfg = make_config(_chain_layers(solver.qnetwork), flux_shape, len(action_map), obs_dtype=solver.obs_dtype,
                      dueling=solver.dueling, double_q=solver.double_q, prioritized_replay=solver.prioritized_replay,
                      batch_size=solver.batch_size, buffer_size=solver.buffer_size, learning_rate=solver.learning_rate,
                      discount=default_discount(env), seed=solver.seed, device=solver.device, math_mode=solver.math_mode,
                      trace_length=solver.trace_length if solver.recurrence else 0,
                      max_episode_length=solver.max_episode_length if solver.recurrence else 0, max_act_rows=max(2 * solver.batch_size, 128))
    engine = Engine(cfg)
    engine.set_params(flat_params(active_q))
    replay = initialize_replay_buffer(solver, env, action_indices, engine)
    policy = NNPolicy(env, active_q, action_map, obs.ndim, engine)
    return dqn_train(solver, env, policy, replay)


def batch_train(solver, env, policy, optimizer, target_q, replay, discount=None):
    """src/solver.jl:191-236 -> (loss_val, grad_norm).  The optimiser state and the target network live in the
    engine; `optimizer` and `target_q` are accepted for signature parity and are not read."""
    return policy.engine.train_step()


def save_model(solver, policy, scores_eval, saved_mean_reward, model_saved):
    """src/solver.jl:290-300 (npz list of the Flux.params arrays instead of BSON)."""
    if scores_eval >= saved_mean_reward:
        os.makedirs(solver.logdir, exist_ok=True)
        np.savez(os.path.join(solver.logdir, "qnetwork.npz"), qnetwork=policy.engine.get_params())
        if solver.verbose:
            print("Saving new model with eval reward %1.3f " % scores_eval)
        model_saved, saved_mean_reward = True, scores_eval
    return model_saved, saved_mean_reward


def dqn_train(solver, env, policy, replay):
    """src/solver.jl:59-178, same cadence: train every train_freq env steps, hard target sync every
    target_update_freq env steps."""
    engine = policy.engine
    engine.sync_target()                                       # target_q = deepcopy(active_q)  :65
    optimizer = target_q = None                                # live in the engine
    resetstate(policy)
    env.reset()
    obs = env.observe()
    step = 0
    episode_rewards, episode_steps = [0.0], []
    saved_mean_reward, scores_eval = -np.inf, -np.inf
    model_saved = eval_next = save_next = False
    loss_val = grad_val = float("nan")
    action_indices = {a: i + 1 for i, a in enumerate(policy.action_map)}
    for t in range(1, solver.max_steps + 1):
        act = solver.exploration_policy.action(policy, t, obs)
        ai = action_indices[act]
        rew = env.act(act)
        op = env.observe()
        done = env.terminated()
        exp = DQExperience(np.asarray(obs), ai, np.float32(rew), np.asarray(op), done)
        if solver.recurrence:
            replay.add_exp(exp)                                                             # :89-90
        else:
            replay.add_exp(exp, abs(exp.r) if solver.prioritized_replay else np.float32(0))     # :91-95
        obs = op
        step += 1
        episode_rewards[-1] += rew
        if done or step >= solver.max_episode_length:
            if eval_next:
                scores_eval, steps_eval, info_eval = solver.evaluation_policy(policy, env, solver.num_ep_eval, solver.max_episode_length, solver.verbose)
                eval_next = False
                if save_next and solver.logdir is not None:
                    model_saved, saved_mean_reward = save_model(solver, policy, scores_eval, saved_mean_reward, model_saved)
                    save_next = False
                solver.history.append(dict(t=t, eval_reward=scores_eval, eval_steps=steps_eval))
            env.reset()
            obs = env.observe()
            resetstate(policy)
            episode_steps.append(step)
            episode_rewards.append(0.0)
            done = False
            step = 0
        avg100_reward = float(np.mean(episode_rewards[max(0, len(episode_rewards) - 102):]))
        if t % solver.train_freq == 0:
            loss_val, grad_val = batch_train(solver, env, policy, optimizer, target_q, replay)
        if t % solver.target_update_freq == 0:
            engine.sync_target()                               # Flux.loadparams!(target_q, params(active_q)) :142-145
        if t % solver.eval_freq == 0:
            eval_next = True
        if t % solver.save_freq == 0:
            save_next = True
        if t % solver.log_freq == 0:
            nt = solver.exploration_policy.loginfo(t)
            solver.history.append(dict(t=t, avg_reward=avg100_reward, loss=loss_val, grad_val=grad_val, **nt))
            if solver.verbose:
                print("%5d / %5d eps %0.3f |  avgR %1.3f | Loss %2.3e | Grad %2.3e | EvalR %1.3f " %
                      (t, solver.max_steps, list(nt.values())[0], avg100_reward, loss_val, grad_val, scores_eval))
    if model_saved and solver.verbose:                         # quirk preserved: restore only when verbose (:170-176)
        print("Restore model with eval reward %1.3f " % saved_mean_reward)
        engine.set_params(np.load(os.path.join(solver.logdir, "qnetwork.npz"))["qnetwork"])
    load_flat_params(policy.qnetwork, engine.get_params())     # hand the trained weights back to the host descriptor
    return policy
