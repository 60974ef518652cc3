"""ELL_SNG table on the B200 against the same kernel source on the host (tests/host emulator), point by
point, for one radius: which points flip between "collapses" and "never collapses", by symmetry class
(ix == 0: l1 == l2;  iy == 0: l2 == l3).  Diagnosis of the round-1 hardware failure of
tests/test_zgpu_5_collapse_tables.py::test_ell_sng_tables.   python scripts/gpu_ctable_diag.py [ismooth]"""
import ctypes
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from emu_util import PD, load_emulator, ptr  # noqa: E402
from oracle import pinocchio_oracle as po  # noqa: E402
from pinocchio_b200.cosmology import Cosmology  # noqa: E402
from pinocchio_b200.engine import CT_SNG, Pinocchio, RunConfig  # noqa: E402

ism = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N = 32
ND, NXY = po.CT_NBINS_D, po.CT_NBINS_XY
c = Cosmology()
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7), c)
pin.initialize_collapse_times(CT_SNG)
gpu = pin.collapse_table(ism).ravel()
lib = load_emulator()
lib.emu_ct_build.argtypes = [ctypes.c_int, PD, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, PD, ctypes.c_int,
                             ctypes.c_double, PD, ctypes.c_int, ctypes.c_int, PD]
c4 = np.array([c.p.Omega0, c.p.OmegaLambda, c.OmegaRad, c.OmegaK, 0.0, po.H_OVER_C, 1.0])
D_in = c.GrowingMode(1.0 / 1.0e-5 - 1.0)
dv = po.ct_delta_vector()
cpu = np.zeros(ND * NXY * NXY)
ampl = float(np.sqrt(pin.Smoothing.Variance[ism]))
assert lib.emu_ct_build(3, ptr(dv), ND, NXY, 3.5 / NXY, ampl, None, 0, D_in, ptr(c4), 0, cpu.size, ptr(cpu)) == 0
gold = np.load(ROOT / "tests" / "golden" / "reference_ct_32.npz")
print(f"radius {ism}: nonzero gpu {(gpu != 0).sum()}  host emulator {(cpu != 0).sum()}  reference program {gold['sng_nonzero_per_radius'][ism]}")
i = np.arange(cpu.size)
ix, iy = (i // ND) % NXY, i // ND // NXY
flip = (gpu != 0) != (cpu != 0)
for name, m in (("ix==0 (l1==l2)", ix == 0), ("iy==0 (l2==l3)", (iy == 0) & (ix != 0)), ("generic", (ix != 0) & (iy != 0))):
    print(f"  {name:16s}: points {m.sum():6d}  nonzero(host) {((cpu != 0) & m).sum():6d}  flips {(flip & m).sum():5d}")
both = (gpu != 0) & (cpu != 0)
rel = np.abs(gpu - cpu)[both] / np.maximum(cpu[both], 1e-3)
print(f"  both nonzero: max rel diff {rel.max():.3e}, > 1e-7: {(rel > 1e-7).sum()}")
for k in np.nonzero(flip)[0][:8]:
    print(f"    flip at i={k} id={k % ND} ix={ix[k]} iy={iy[k]} gpu={gpu[k]:.6g} host={cpu[k]:.6g}")
pin.close()
