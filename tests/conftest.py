import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


def _cuda_device_present() -> bool:
    """True when pinb200_create can find an sm_100 device (asked of the driver, not of torch)."""
    import ctypes
    for name in ("libcuda.so.1", "libcuda.so"):
        try:
            cu = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        if cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0:
            return True
    return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped on a box without a CUDA device, so that a plain `pytest tests` is green
    on the development machine; on the B200 box nothing is skipped.  PINB200_LIB (dry runs of the GPU
    tests over the emulated ABI) keeps them runnable."""
    import os
    import pytest
    if os.environ.get("PINB200_LIB") or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a): run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
