"""CPU oracle for the DQN minibatch-update path (TEST INFRASTRUCTURE, not product code).

This package is a CPU restatement of the `batch_train!` hot path of
JuliaPOMDP/DeepQLearning.jl (reference citations are into /root/reference):

  src/solver.jl:191-236                        batch_train! (PER / feed-forward)
  src/prioritized_experience_replay.jl:39-104  PrioritizedReplayBuffer, add_exp!, update_priorities!,
                                               sample, get_batch
  src/dueling.jl:8-11, 36-58                   DuelingNetwork forward, create_dueling_network
  src/helpers.jl:6-19, 38-46                   flattenbatch, huber_loss, globalnorm
  src/solver.jl:239-287, src/episode_replay.jl  recurrent batch_train!, EpisodeReplayBuffer (oracle/recurrent.py)

PARITY UNPINNED: the reference is pure Julia, Julia is not installed in this image, the
arithmetic lives in un-vendored third-party packages (Flux 0.14 / Zygote / NNlib / StatsBase
0.32-0.34, no Manifest.toml), and the reference's own tests hold no golden vector, known-answer
test or fixture for this path (test/runtests.jl asserts only average-return thresholds and
output shapes).  The oracle is therefore pinned by (1) an independent derivation of every
gradient through torch-CPU autograd, (2) closed-form checks, (3) an fp64 evaluation of the same
step, and (4) committed golden vectors minted from this restatement (tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / CPU baseline - never as a product path.
"""
from .philox import philox4x32, uniform24
from .sumtree import SumTree
from .nets import (ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID, Conv, Dense, Flatten,
                   create_dueling_network, DuelingNetwork, Chain, glorot_uniform_chain,
                   params_of, set_params, flat_params, num_params)
from .replay import PrioritizedReplayBuffer, DQExperience, pairwise_sum_f32, pow_f32
from .step import batch_train, Adam, huber_loss, globalnorm, q_targets_of, forward_backward
from .recurrent import (LSTM, RecurrentQ, make_recurrent_q, EpisodeReplayBuffer, forward_backward_recurrent, batch_train_recurrent)
