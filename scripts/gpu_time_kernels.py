"""Per-kernel device times of the radius loop and the displacement stage from the engine's own
CUDA-event timers (no torch): python scripts/gpu_time_kernels.py [N] [steps]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pinocchio_b200.cosmology import Cosmology, SmoothingLadder  # noqa: E402
from pinocchio_b200.engine import Pinocchio, RunConfig  # noqa: E402

HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.time()
cosmo = Cosmology(pk_norm_override=2.03146e7)
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo,
                smoothing=SmoothingLadder(np.array(HMF_RADII), np.zeros(9)))
pin.GenIC_large()
if steps == 0:                           # under ncu: one sweep, no displacement stage
    pin.compute_fmax(displacements=False)
    pin.close()
    sys.exit(0)
pin.compute_fmax()                       # warm-up
a = pin.timers()
w0 = time.perf_counter()
for _ in range(steps):
    pin.compute_fmax()
wall = (time.perf_counter() - w0) / steps
b = pin.timers()
S = len(HMF_RADII)
out = {"grid": N, "steps": steps, "wall_ms_per_step": round(wall * 1e3, 2),
       "xpass_ms": round((b.hess_x - a.hess_x) / (steps * S) * 1e3, 3),
       "ypass_ms": round((b.hess_y - a.hess_y) / (steps * S) * 1e3, 3),
       "zpass_collapse_ms": round((b.hess_z - a.hess_z) / (steps * S) * 1e3, 3),
       "fmax_ms": round((b.fmax - a.fmax) / steps * 1e3, 2), "lpt_ms": round((b.lpt - a.lpt) / steps * 1e3, 2),
       "sigma_R0": round(float(np.sqrt(pin.TrueVariance[-1])), 4), "pdf_total_ok": bool(int(pin.Fmax_PDF().sum()) == N ** 3),
       "setup_s": round(time.time() - t0 - wall * steps, 1)}
print(json.dumps(out))
pin.close()
