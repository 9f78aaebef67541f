"""ctypes binding of include/dqn_b200.h - the same C-ABI the Julia host binds with ccall.

There is no CPU fallback: if libdqn_b200.so is missing or fails to load, importing this module raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdqn_b200.so")

DQN_ABI_VERSION = 1
DQN_MAX_LAYERS = 16
DQN_NCCL_ID_BYTES = 128
DQN_OK, DQN_ERR_INVALID, DQN_ERR_CUDA, DQN_ERR_STATE, DQN_ERR_NCCL, DQN_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
LAYER_DENSE, LAYER_CONV, LAYER_FLATTEN, LAYER_LSTM = 0, 1, 2, 3
OBS_F32, OBS_U8 = 0, 1
NET_ONLINE, NET_TARGET = 0, 1
Q_S_ONLINE, Q_SP_ONLINE, Q_SP_TARGET = 0, 1, 2
MATH_FP32, MATH_3XTF32 = 0, 1


class dqn_layer_t(C.Structure):
    _fields_ = [("kind", C.c_int32), ("act", C.c_int32), ("in_", C.c_int32), ("out", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32)]


class dqn_config_t(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32),
                ("obs_c", C.c_int32), ("obs_h", C.c_int32), ("obs_w", C.c_int32), ("obs_dtype", C.c_int32),
                ("n_actions", C.c_int32), ("n_layers", C.c_int32), ("layers", dqn_layer_t * DQN_MAX_LAYERS),
                ("dueling", C.c_int32), ("double_q", C.c_int32), ("prioritized_replay", C.c_int32), ("batch_size", C.c_int32),
                ("buffer_size", C.c_int64), ("alpha", C.c_float), ("beta", C.c_float), ("eps", C.c_float),
                ("learning_rate", C.c_float), ("discount", C.c_float),
                ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
                ("seed", C.c_uint64), ("math_mode", C.c_int32), ("use_graph", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("nccl_id", C.c_uint8 * DQN_NCCL_ID_BYTES), ("max_act_rows", C.c_int32), ("trace_length", C.c_int32), ("max_episode_length", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class DQNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdqn_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing - build it with `python deepqlearning.jl_b200/build.py` "
                      "(there is no CPU fallback for the DQN step)")
lib = C.CDLL(LIB_PATH)

_H = C.c_void_p
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); must list every symbol include/dqn_b200.h declares (tests/test_capi_cpu.py checks)
SIGNATURES = {
    "dqn_config_default": (C.c_int, [C.POINTER(dqn_config_t)]),
    "dqn_engine_create": (C.c_int, [C.POINTER(dqn_config_t), C.POINTER(_H)]),
    "dqn_engine_destroy": (None, [_H]),
    "dqn_last_error": (C.c_char_p, [_H]),
    "dqn_nccl_unique_id": (C.c_int, [_u8p]),
    "dqn_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dqn_group_create": (C.c_int, [C.POINTER(dqn_config_t), C.c_int, C.POINTER(C.c_int), C.POINTER(_H)]),
    "dqn_group_destroy": (None, [_H]),
    "dqn_group_size": (C.c_int, [_H]),
    "dqn_group_engine": (_H, [_H, C.c_int]),
    "dqn_group_last_error": (C.c_char_p, [_H]),
    "dqn_group_set_params": (C.c_int, [_H, C.c_int, _f32p, C.c_int64]),
    "dqn_group_sync_target": (C.c_int, [_H]),
    "dqn_group_train_step": (C.c_int, [_H, _f32p, _f32p]),
    "dqn_num_params": (C.c_int64, [_H]),
    "dqn_set_params": (C.c_int, [_H, C.c_int, _f32p, C.c_int64]),
    "dqn_get_params": (C.c_int, [_H, C.c_int, _f32p, C.c_int64]),
    "dqn_sync_target": (C.c_int, [_H]),
    "dqn_get_adam_state": (C.c_int, [_H, _f32p, _f32p, C.POINTER(C.c_double), C.c_int64]),
    "dqn_replay_add": (C.c_int, [_H, C.c_void_p, _i32p, _f32p, C.c_void_p, _u8p, _f32p, C.c_int64]),
    "dqn_replay_add_device": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "dqn_replay_size": (C.c_int, [_H, _i64p, _i64p]),
    "dqn_replay_fill_synthetic": (C.c_int, [_H, C.c_int64, C.c_uint64]),
    "dqn_replay_read": (C.c_int, [_H, _i64p, C.c_int64, C.c_void_p, _i32p, _f32p, C.c_void_p, _u8p]),
    "dqn_episode_add": (C.c_int, [_H, _f32p, _i32p, _f32p, _f32p, _u8p, C.c_int64]),
    "dqn_episode_count": (C.c_int, [_H, _i64p, _i64p]),
    "dqn_episode_sample": (C.c_int, [_H, C.c_uint64, _i64p, _i32p]),
    "dqn_policy_reset": (C.c_int, [_H]),
    "dqn_update_priorities": (C.c_int, [_H, _i64p, _f32p, C.c_int64]),
    "dqn_set_priorities": (C.c_int, [_H, _i64p, _f32p, C.c_int64]),
    "dqn_get_priorities": (C.c_int, [_H, _f32p, C.c_int64]),
    "dqn_get_tree": (C.c_int, [_H, _f32p, C.c_int64]),
    "dqn_tree_nodes": (C.c_int64, [_H]),
    "dqn_sample_indices": (C.c_int, [_H, C.c_uint64, _i64p]),
    "dqn_get_batch": (C.c_int, [_H, _i64p, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p]),
    "dqn_train_step": (C.c_int, [_H, _f32p, _f32p]),
    "dqn_train_step_with_indices": (C.c_int, [_H, _i64p, _f32p, _f32p]),
    "dqn_train_step_async": (C.c_int, [_H]),
    "dqn_sync": (C.c_int, [_H, _f32p, _f32p]),
    "dqn_step_result": (C.c_int, [_H, C.c_int, _f32p, _f32p]),
    "dqn_q_values": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64, _f32p]),
    "dqn_act": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_float, C.c_uint64, _i32p, _f32p]),
    "dqn_act_device": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p]),
    "dqn_synth_env_step": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64]),
    "dqn_get_last_indices": (C.c_int, [_H, _i64p]),
    "dqn_get_td": (C.c_int, [_H, _f32p]),
    "dqn_get_is_weights": (C.c_int, [_H, _f32p]),
    "dqn_get_q": (C.c_int, [_H, C.c_int, _f32p]),
    "dqn_get_targets": (C.c_int, [_H, _f32p, _i32p]),
    "dqn_get_grads": (C.c_int, [_H, _f32p, C.c_int64]),
    "dqn_get_activation": (C.c_int, [_H, C.c_int, C.c_int, _f32p, C.c_int64]),
    "dqn_timer_start": (C.c_int, [_H]),
    "dqn_timer_stop": (C.c_int, [_H, _f32p]),
    "dqn_launches_per_step": (C.c_int, [_H]),
    "dqn_collective_kind": (C.c_int, [_H]),
    "dqn_set_profiling": (C.c_int, [_H, C.c_int]),
    "dqn_get_profile": (C.c_int, [_H, C.c_char_p, C.c_int64]),
    "dqn_flush_l2": (C.c_int, [_H]),
    "dqn_stream": (C.c_void_p, [_H]),
    "dqn_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "dqn_host_free": (C.c_int, [C.c_void_p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def ptr(a, ctype):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


def vptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def default_config():
    cfg = dqn_config_t()
    lib.dqn_config_default(C.byref(cfg))
    return cfg


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'd memory (for the end-to-end paths of bench.py)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = lib.dqn_host_alloc(C.byref(p), max(n, 1))
    if rc != 0:
        raise DQNError(rc, "cudaHostAlloc failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr
