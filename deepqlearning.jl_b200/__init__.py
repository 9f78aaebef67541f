"""dqn_b200 - B200-native engine behind the `batch_train!` hot path of JuliaPOMDP/DeepQLearning.jl.

Host mirror of the reference's surface for that path (names follow src/DeepQLearning.jl:19-33):
    DeepQLearningSolver, solve, dqn_train, batch_train, NNPolicy, PrioritizedReplayBuffer, DQExperience,
    DuelingNetwork, create_dueling_network, flattenbatch, Chain, Dense, Conv
All numerical work happens in libdqn_b200.so (hand-written sm_100a CUDA kernels) through the C-ABI of
include/dqn_b200.h; importing this package fails if that library is missing."""
from . import _capi
from ._capi import DQNError, MATH_FP32, MATH_3XTF32
from .engine import Engine, Group, make_config, nccl_unique_id
from .flux import (Chain, Dense, Conv, LSTM, flattenbatch, DuelingNetwork, create_dueling_network, isrecurrent, flat_params,
                   load_flat_params, identity, relu, tanh, sigmoid)
from .dist import ControlPlane, shard_seeds
from .replay import PrioritizedReplayBuffer, EpisodeReplayBuffer, DQExperience
from .solver import (DeepQLearningSolver, solve, dqn_train, batch_train, NNPolicy, EpsGreedyPolicy, LinearDecaySchedule,
                     basic_evaluation, batched_evaluation, getnetwork, actionvalues, action, value, initialize_replay_buffer, populate_replay_buffer)
