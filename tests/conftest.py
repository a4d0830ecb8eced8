import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def m5lib():
    from mars5_tts_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def bare_ctx(m5lib):
    """A context without weights: enough for the kernel-level m5_dbg_* entry points."""
    import ctypes as C
    from mars5_tts_b200 import capi
    cfg = capi.ModelCfg()
    ctx = C.c_void_p()
    rc = m5lib.m5_create(0, C.byref(cfg), None, 0, C.byref(ctx))
    assert rc == 0, f"m5_create failed: {rc}"
    yield ctx
    m5lib.m5_destroy(ctx)
