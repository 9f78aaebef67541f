"""No-GPU checks of the boundary: the library loads, exports every symbol include/dqn_b200.h declares, the
ctypes struct matches the C struct, the operand functors pass their CPU index-algebra test, and the product
path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dqn_b200.h")


def test_library_exports_every_declared_symbol(lib):
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(dqn_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 40
    from dqn_b200 import _capi
    for name in sorted(declared):
        assert hasattr(_capi.lib, name), f"{name} declared in dqn_b200.h but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature"


def test_config_struct_layout_matches_c(lib, tmp_path):
    from dqn_b200 import _capi
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dqn_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(dqn_config_t),'
                    ' offsetof(dqn_config_t, layers), offsetof(dqn_config_t, buffer_size), offsetof(dqn_config_t, seed), offsetof(dqn_config_t, nccl_id));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    T = _capi.dqn_config_t
    assert got == [C.sizeof(T), T.layers.offset, T.buffer_size.offset, T.seed.offset, T.nccl_id.offset]


def test_defaults_are_the_reference_defaults(lib):
    from dqn_b200 import _capi
    cfg = _capi.default_config()
    # src/prioritized_experience_replay.jl:43-45, src/solver.jl:3,5,10,11,16,20
    assert (round(cfg.alpha, 6), round(cfg.beta, 6), round(cfg.eps, 6)) == (0.6, 0.4, 0.001)
    assert (cfg.batch_size, cfg.buffer_size, cfg.dueling, cfg.double_q, cfg.prioritized_replay) == (32, 1000, 1, 1, 1)
    assert abs(cfg.learning_rate - 1e-4) < 1e-10 and (cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps) == (0.9, 0.999, 1e-8)


def test_no_cpu_fallback(lib):
    from dqn_b200 import _capi, Engine, make_config, DQNError
    n = C.c_int(0)
    _capi.lib.dqn_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    cfg = make_config([dict(kind=0, act=0, in_=2, out=4)], (2,), 4)
    with pytest.raises(DQNError) as ei:
        Engine(cfg)
    assert ei.value.code in (_capi.DQN_ERR_CUDA,)


def test_bad_topology_is_an_error_code_not_a_crash(lib):
    from dqn_b200 import _capi
    cfg = _capi.default_config()
    cfg.abi_version = 99
    h = C.c_void_p()
    assert _capi.lib.dqn_engine_create(C.byref(cfg), C.byref(h)) == _capi.DQN_ERR_INVALID
    assert b"abi_version" in _capi.lib.dqn_last_error(None)


def test_operand_functors_on_cpu(tmp_path):
    exe = tmp_path / "ops"
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O1", "-std=c++17", "-w", "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "test_ops_host.cu")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout
