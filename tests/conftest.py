import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    """True when librt_b200.so loads and rt_create finds a CUDA device (no torch needed for the check)."""
    try:
        import ctypes as C
        from raytracer_js_b200 import _native as N
        lib = N.load()
        ctx = C.c_void_p()
        if lib.rt_create(-1, C.byref(ctx)) != N.RT_OK:
            return False
        lib.rt_destroy(ctx)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a box without a GPU: the gpu-marked tests are skipped, not failed.  With `-m gpu` they
    run regardless - on the GPU box a missing device or library must FAIL loudly, never skip."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the B200 box)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    orc.lib()
    return orc
