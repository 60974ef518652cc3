"""Times the device-side fragmentation hand-off (pinb200_collapsed_cells + sorted AoS download) on one
GPU and checks its result: python scripts/gpu_handoff_probe.py [N].  Prints one JSON line.
Run by bench.py in a fresh process after its own measurements (no torch here)."""
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
cosmo = Cosmology(pk_norm_override=2.03146e7)
pin = Pinocchio(RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3), cosmo,
                smoothing=SmoothingLadder(np.array(HMF_RADII), np.zeros(9)))
pin.GenIC_large()
pin.compute_fmax()
t0 = time.perf_counter()
idx = pin.collapsed_cells(1.0)          # count query, then selection + radix sort + D2H of the index list into pageable memory
t_sort = time.perf_counter() - t0
t_dev = pin.timers().sort_ms            # selection + sort on the device clock (CUDA events)
F = pin.field("Fmax").ravel()
Fs = F[idx]
ok = bool(idx.size == int((F >= 1.0).sum()) and (np.diff(Fs) <= 0).all())
chunk = min(idx.size, 1 << 24)
t0 = time.perf_counter()
frag = pin.sorted_products(0, chunk)    # gathered AoS records, frag[] order
t_dl = time.perf_counter() - t0
ok = ok and bool(np.array_equal(frag["Fmax"], Fs[:chunk]))
print(json.dumps({"grid": N, "collapsed_cells": int(idx.size), "collapsed_fraction": round(idx.size / float(N) ** 3, 6),
                  "filter_sort_ms": round(t_dev, 2), "filter_sort_wall_ms_incl_pageable_index_download": round(t_sort * 1e3, 2), "sorted_records_downloaded": int(chunk),
                  "sorted_download_ms": round(t_dl * 1e3, 2), "order_and_records_ok": ok}))
pin.close()
