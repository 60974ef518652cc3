"""The `-m gpu` tests that round 1 could not run on hardware, executed here against the test-only
emulated ABI (oracle/_ref/libpinb_emuabi.so = tests/host/emu_abi.cpp: the same C ABI, the same kernel
bodies under the CPU block emulator, grids 32 and 64): checks the tests' own logic -- expected orders,
record gathers, tolerances, error paths -- so that a failure on the B200 points at the device code.
The library is selected with PINB200_LIB (pinocchio_b200/engine.py); each file runs in its own pytest
process.  This is not a CPU fallback of the product: libpinb200.so has none (tests/test_abi.py).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EMU_ABI = ROOT / "oracle" / "_ref" / "libpinb_emuabi.so"

pytestmark = pytest.mark.skipif(not EMU_ABI.exists(), reason="oracle/_ref/libpinb_emuabi.so not built (make -C oracle all)")


@pytest.mark.parametrize("target,select,npass", [
    # fragmentation hand-off: the 64^3 case (128^3 and 512^3 are beyond the emulated ABI)
    ("tests/test_zgpu_3_fragment_handoff.py", "records and 64", 1),
    # collapse-time tables without the linked programs (those run in test_collapse_tables.py)
    ("tests/test_zgpu_5_collapse_tables.py", "not linked", 4),
    # -DDOUBLE_PRECISION_PRODUCTS records and lpt_order 1 / 2 through the ABI (the linked programs run in test_dropin_emulated.py)
    ("tests/test_zgpu_6_build_variants.py", "widened or lower_lpt or seed_plane", 4),
    # snapshot blocks and product dumps written from the device SoA: the 32^3 case
    ("tests/test_zgpu_7_device_writers.py", "from_device and 32", 1),
])
def test_late_gpu_tests_pass_on_the_emulated_abi(target, select, npass):
    env = dict(os.environ, PINB200_LIB=str(EMU_ABI))
    r = subprocess.run([sys.executable, "-m", "pytest", target, "-m", "gpu", "-q", "-x", "-k", select, "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert f"{npass} passed" in r.stdout, tail
