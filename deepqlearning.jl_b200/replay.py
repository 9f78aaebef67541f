"""PrioritizedReplayBuffer with the reference's surface, HBM-resident (src/prioritized_experience_replay.jl).

    DQExperience(s, a, r, sp, done)            :3-9     a is 1-based, as in the reference
    add_exp!(r, expe, td_err=abs(expe.r))      :65-74   -> add_exp
    update_priorities!(r, indices, td_errors)  :76-80   -> update_priorities
    StatsBase.sample(r)                        :82-87   -> sample   (sum-tree instead of the O(N) scan)
    get_batch(r, sample_indices)               :89-104  -> get_batch

`sample`/`get_batch` copy a batch back to the host for callers that want to look at it; the training step
itself (`batch_train`) never does - it samples and gathers on the device."""
from dataclasses import dataclass

import numpy as np


@dataclass
class DQExperience:
    s: np.ndarray
    a: int
    r: float
    sp: np.ndarray
    done: bool


class PrioritizedReplayBuffer:
    def __init__(self, engine):
        self.engine = engine
        self.max_size = int(engine.cfg.buffer_size)
        self.batch_size = int(engine.cfg.batch_size)
        self.alpha, self.beta, self.eps = engine.cfg.alpha, engine.cfg.beta, engine.cfg.eps
        self._sample_calls = 0

    @property
    def _curr_size(self):
        return self.engine.replay_size()[0]

    @property
    def _idx(self):               # 1-based like the reference's r._idx
        return self.engine.replay_size()[1] + 1

    def is_full(self):
        return self._curr_size == self.max_size

    def add_exp(self, expe, td_err=None):
        td = abs(np.float32(expe.r)) if td_err is None else td_err
        self.engine.replay_add(expe.s[None], [expe.a], [expe.r], expe.sp[None], [expe.done], [td])

    def add_batch(self, s, a, r, sp, done, td_err):
        self.engine.replay_add(s, a, r, sp, done, td_err)

    def update_priorities(self, indices, td_errors):
        self.engine.update_priorities(indices, td_errors)

    def sample_indices(self, call=None):
        if call is None:                      # the reference's sample(r) advances the buffer's RNG on every call (PER:82-87)
            call = self._sample_calls
            self._sample_calls += 1
        return self.engine.sample_indices(call)

    def sample(self):
        if self._curr_size < self.batch_size:
            raise AssertionError("r._curr_size >= r.batch_size")
        return self.get_batch(self.sample_indices())

    def get_batch(self, sample_indices):
        return self.engine.get_batch(sample_indices)

    @property
    def _priorities(self):
        return self.engine.get_priorities()


class EpisodeReplayBuffer:
    """src/episode_replay.jl:3-95 with the episodes resident in HBM.  add_exp! collects the running episode on the host and hands it
    to the engine when it ends (:52-58); `max_size` counts episodes.  One guard the reference lacks: a running episode that reaches the
    engine's max_episode_length without `done` is stored as it is (the reference keeps appending across resets)."""

    def __init__(self, engine):
        self.engine = engine
        self.max_size = int(engine.cfg.buffer_size)
        self.batch_size = int(engine.cfg.batch_size)
        self.trace_length = int(engine.T)
        self.max_len = int(engine.cfg.max_episode_length or 100)
        self._episode = []

    @property
    def _curr_size(self):
        return self.engine.episode_count()[0]

    def is_full(self):
        return self._curr_size == self.max_size

    def add_exp(self, expe, td_err=None):
        self._episode.append(expe)
        if expe.done or len(self._episode) >= self.max_len:
            self.add_episode(self._episode)
            self._episode = []

    def add_episode(self, ep):
        self.engine.episode_add(np.stack([np.asarray(e.s, np.float32) for e in ep]), [e.a for e in ep], [e.r for e in ep],
                                np.stack([np.asarray(e.sp, np.float32) for e in ep]), [e.done for e in ep])
