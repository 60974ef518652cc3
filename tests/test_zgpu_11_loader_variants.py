"""The strided passes have two tile loaders (TMA: cp.async.bulk.tensor + mbarrier; LDGSTS: cp.async) and two
schedules (blocks that walk the tiles with the next tile's load under this tile's stores; one tile per block),
selected at run time by PINB200_TMA / PINB200_PERSISTENT (k_strided.cu).  All four combinations move the same
bytes through the same butterflies: the products of a 64^3 run must be bit-identical, on the product build and on
the split build (decimated long lines).  One fresh process per combination (the switches are read once).  -m gpu."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SPLIT = ROOT / "pinocchio_b200" / "libpinb200_split.so"
pytestmark = pytest.mark.gpu

SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
from pinocchio_b200.engine import Pinocchio, RunConfig
N = 64
cosmo = Cosmology(pk_norm_override=2.03146e7)
lad = SmoothingLadder(np.array([9.026099, 3.058354, 0.689079, 0.0]), np.zeros(4))
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo, smoothing=lad)
pin.GenIC_large()
pin.compute_fmax()
h = hashlib.sha256()
h.update(pin.field("Fmax").tobytes()); h.update(pin.field("Rmax").tobytes())
for n in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"):
    for a in range(3):
        h.update(pin.field(n, a).tobytes())
for w in range(3):
    h.update(pin.read_kvector(w).tobytes())
print("DIGEST", h.hexdigest(), float(pin.field("Fmax").max()))
pin.close()
""" % str(ROOT)


def digest(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1].split()
    assert float(line[2]) > 1.0          # a real field: some cell collapses
    return line[1]


@pytest.mark.parametrize("lib", ["product", "split"])
def test_loaders_and_schedules_are_bit_identical(lib):
    base = {}
    if lib == "split":
        if not SPLIT.exists():
            pytest.skip("libpinb200_split.so not built")
        base = {"PINB200_LIB": str(SPLIT)}
    ref = digest({**base, "PINB200_TMA": "0", "PINB200_PERSISTENT": "0"})       # the round-1 loader and schedule
    for tma, pers in (("1", "1"), ("1", "0"), ("0", "1")):
        assert digest({**base, "PINB200_TMA": tma, "PINB200_PERSISTENT": pers}) == ref, (lib, tma, pers)
