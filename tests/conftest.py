import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count():
    """CUDA devices as libdqn_b200 itself sees them (0 on a CPU-only host, or when the library cannot be loaded)."""
    try:
        import ctypes as C
        import dqn_b200
        n = C.c_int(0)
        return n.value if dqn_b200._capi.lib.dqn_device_count(C.byref(n)) == 0 else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # a host without a GPU skips the `gpu` tests instead of failing at dqn_engine_create (the library has no CPU path)
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: libdqn_b200 has no CPU path")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    import dqn_b200
    return dqn_b200


@pytest.fixture(scope="session")
def n_gpus():
    return _device_count()
