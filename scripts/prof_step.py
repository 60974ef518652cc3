"""One short run of the hot path for profilers: GenIC, then `reps` sweeps of a `nrad`-radius ladder (+ 3LPT).
    python scripts/prof_step.py [N=1024] [classic|tab] [nrad=2] [reps=2] [lpt=1]
Prints the engine's per-kernel CUDA-event timers of the LAST sweep as one JSON line."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pinocchio_b200.cosmology import Cosmology, SmoothingLadder, set_smoothing  # noqa: E402
from pinocchio_b200.engine import CT_CLASSIC, Pinocchio, RunConfig  # noqa: E402

HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "classic"
nrad = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
lpt = int(sys.argv[5]) if len(sys.argv) > 5 else 1
cosmo = Cosmology(pk_norm_override=2.03146e7)
var = set_smoothing(cosmo, 1.0 / 0.7).Variance
sel = list(range(9))[-nrad:] if nrad < 9 else list(range(9))
if nrad == 2:
    sel = [4, 8]
lad = SmoothingLadder(np.array([HMF_RADII[i] for i in sel]), np.array([var[i] for i in sel]))
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo, smoothing=lad)
pin.GenIC_large()
if mode == "tab":
    pin.initialize_collapse_times(CT_CLASSIC)
for _ in range(reps):
    t0 = pin.timers()
    pin.compute_fmax(displacements=bool(lpt))
    t1 = pin.timers()
S = len(sel)
print(json.dumps({"grid": N, "mode": mode, "radii": S, "x_ms": round((t1.hess_x - t0.hess_x) * 1e3 / S, 3),
                  "y_ms": round((t1.hess_y - t0.hess_y) * 1e3 / S, 3), "z_ms": round((t1.hess_z - t0.hess_z) * 1e3 / S, 3),
                  "lpt_ms": round((t1.lpt - t0.lpt) * 1e3, 3), "pdf_total_ok": bool(int(pin.Fmax_PDF().sum()) == N ** 3)}))
pin.close()
