"""Times the collapse-time tables (-DTABULATED_CT / -DELL_SNG) on one GPU:
python scripts/gpu_ctable_probe.py [N].  Prints one JSON line.  Table build for the nine radii with
ell_classic and with the ELL_SNG batch ODE kernel (250 000 rkf45 integrations per radius), then the Fmax
sweep with the table look-up in the collapse z pass next to the direct ell_classic sweep.
Run by bench.py in a fresh process after its own measurements (no torch here)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pinocchio_b200.cosmology import Cosmology, SmoothingLadder, set_smoothing  # noqa: E402
from pinocchio_b200.engine import CT_CLASSIC, CT_SNG, Pinocchio, RunConfig  # noqa: E402

HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cosmo = Cosmology(pk_norm_override=2.03146e7)
var = set_smoothing(cosmo, 1.0 / 0.7).Variance            # 1 Mpc/h cells: the same nine-radius ladder
assert var.size == 9
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo,
                smoothing=SmoothingLadder(np.array(HMF_RADII), var))
pin.GenIC_large()
out = {"grid": N, "table_points_per_radius": 250000, "radii": 9}


def sweep():
    t0 = pin.timers()
    w = time.perf_counter()
    pin.compute_fmax(displacements=False)
    w = time.perf_counter() - w
    t1 = pin.timers()
    return {"wall_ms": round(w * 1e3, 2), "fmax_ms": round((t1.fmax - t0.fmax) * 1e3, 2),
            "zpass_collapse_ms_per_radius": round((t1.hess_z - t0.hess_z) * 1e3 / 9, 3)}


sweep()                                                    # warm-up (module load, allocations)
out["direct_ell_classic"] = sweep()
Fd = pin.field("Fmax").ravel()[::61].astype(np.float64)      # a strided sample of the cells
for name, model in (("classic", CT_CLASSIC), ("sng", CT_SNG)):
    c0 = pin.timers().coll
    w = time.perf_counter()
    pin.initialize_collapse_times(model)
    w = time.perf_counter() - w
    out[f"table_build_{name}"] = {"wall_ms": round(w * 1e3, 2), "device_ms_per_radius": round((pin.timers().coll - c0) * 1e3 / 9, 3)}
    t = pin.collapse_table(5)
    out[f"table_build_{name}"]["nonzero_points_radius5"] = int((t != 0).sum())
    sweep()
    out[f"tabulated_{name}"] = sweep()
    Ft = pin.field("Fmax").ravel()[::61].astype(np.float64)
    m = Fd > 1.0
    out[f"tabulated_{name}"]["median_rel_diff_to_direct"] = float(np.median(np.abs(Ft[m] - Fd[m]) / Fd[m]))
    out[f"tabulated_{name}"]["pdf_total_ok"] = bool(int(pin.Fmax_PDF().sum()) == N ** 3)
print(json.dumps(out))
pin.close()
