"""The code paths the product only takes for the 2048^3 box on eight GPUs -- the decimation-in-frequency strided
passes (XCfg::SPLIT / YCfg::SPLIT, kernels.cuh) and the collapse kernel that reads its spline from global memory
(CollapseCfg::SPLINE_GLOBAL, k_zpass.cu) -- on one B200 at 64^3 / 128^3 against the oracle: the parity tests of
tests/test_gpu_parity.py run again in a fresh process on pinocchio_b200/libpinb200_split.so, the test-only build
of the same sources with the thresholds lowered (pinocchio_b200/build.py VARIANTS["split"], built by
__graft_entry__.build()).  Needs a B200: -m gpu.
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SPLIT = ROOT / "pinocchio_b200" / "libpinb200_split.so"
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not SPLIT.exists(), reason="libpinb200_split.so not built (python -m pinocchio_b200.build --variants)")]


@pytest.mark.parametrize("select,npass", [
    # r2c / c2r through the split x and y passes, 32^3 .. 256^3 (incl. the non-Hermitian c2r semantics)
    ("fft_forward_reverse", 4),
    # Hessians, the whole sweep + 3LPT against the oracle at 64^3 and 128^3, the reference-compiled golden at 32^3
    ("second_derivatives or (fmax_and_displacements and not 256) or against_reference_code_golden", 4),
    # k-dependent growth in the split x-pass loader, RECOMPUTE re-entry
    ("recompute_displacements_reentry", 1),
])
def test_parity_suite_on_the_split_build(select, npass):
    env = dict(os.environ, PINB200_LIB=str(SPLIT))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-k", select,
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert f"{npass} passed" in r.stdout, tail


def test_scale_dependent_growth_on_the_split_build():
    env = dict(os.environ, PINB200_LIB=str(SPLIT))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_zgpu_1_scaledep.py", "-m", "gpu", "-q", "-x", "-k", "against_oracle",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "2 passed" in r.stdout, (r.stdout + r.stderr)[-2000:]
