"""Host side of the drop-in shim (shim/fmax_b200.c) against the reference functions it replaces.

oracle/_ref/libshim_host.so = the shim + the reference's src/variables.c, linked against
libpinb200.so exactly as a PINOCCHIO build would (oracle/Makefile); oracle/_ref/libpinocchio_ref.so
holds the reference's own dump_products / read_dumps / set_one_grid.  Checked here without a GPU:
the DumpProducts file boundary in both directions (src/fmax.c:372-506) and
the slab geometry (src/fmax-pfft.c:80-134).  Each library keeps its own copy of the reference
globals (RTLD_LOCAL), and both need a fresh process per grid, hence the subprocesses.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_runner as rr  # noqa: E402

pytestmark = pytest.mark.skipif(not (rr.available() and rr.SHIM_HOST_LIB.exists()),
                                reason="oracle/_ref not built (needs /root/reference once: make -C oracle all)")

SHIM_SIDE = r"""
import ctypes, sys, numpy as np
lib = ctypes.CDLL(sys.argv[1])
mode, N, d, seed = sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
assert lib.host_setup(N, 0, 1, seed, 9, d.encode()) == 0
lib.host_local_cells.restype = ctypes.c_long
n, sz = lib.host_local_cells(), lib.host_sizeof_product()
assert n == N ** 3 and sz == 56
rng = np.random.default_rng(seed)
rec = rng.integers(0, 256, n * sz, dtype=np.uint8)
tv = rng.standard_normal(9)
P = ctypes.c_void_p
PD = ctypes.POINTER(ctypes.c_double)
if mode == "dump":
    lib.host_set_products(rec.ctypes.data_as(P), tv.ctypes.data_as(PD))
    assert lib.host_dump_products() == 0
else:
    rc = lib.host_read_dumps()
    if mode == "read_expect_fail":
        sys.exit(0 if rc != 0 else 1)
    assert rc == 0
    got, gtv = np.zeros_like(rec), np.zeros(9)
    lib.host_get_products(got.ctypes.data_as(P), gtv.ctypes.data_as(PD))
    assert np.array_equal(got, rec) and np.array_equal(gtv, tv)
"""

REF_SIDE = r"""
import ctypes, sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle.reference_runner import ReferenceRun, HMF_RADII
mode, N, d, seed = sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
x = np.linspace(-2, 0, 8)
run = ReferenceRun(N, float(N), HMF_RADII, np.ones(4), x, x, threads=1)
lib = run.lib
lib.ref_dump_setup(d.encode(), seed)
rng = np.random.default_rng(seed)
rec = rng.integers(0, 256, N ** 3 * 56, dtype=np.uint8)
tv = rng.standard_normal(9)
P = ctypes.c_void_p
PD = ctypes.POINTER(ctypes.c_double)
lib.ref_set_products.argtypes = [P, PD]
lib.ref_read_dumps.argtypes = [PD]
if mode == "dump":
    lib.ref_set_products(rec.ctypes.data_as(P), tv.ctypes.data_as(PD))
    assert run._in_workdir(lambda: lib.ref_dump_products()) == 0
else:
    gtv = np.zeros(9)
    assert run._in_workdir(lambda: lib.ref_read_dumps(gtv.ctypes.data_as(PD))) == 0
    got = np.zeros_like(rec)
    lib.ref_fetch_products(got.ctypes.data_as(P))
    assert np.array_equal(got, rec) and np.array_equal(gtv, tv)
"""


def run_shim(mode, N, d, seed):
    return subprocess.run([sys.executable, "-c", SHIM_SIDE, str(rr.SHIM_HOST_LIB), mode, str(N), str(d), str(seed)],
                          capture_output=True, text=True)


def run_ref(mode, N, d, seed):
    return subprocess.run([sys.executable, "-c", REF_SIDE, str(ROOT), mode, str(N), str(d), str(seed)],
                          capture_output=True, text=True)


def test_dump_products_files_identical_and_cross_readable(tmp_path):
    N, seed = 16, 486604
    a, b = str(tmp_path / "shim") + "/", str(tmp_path / "ref") + "/"
    r = run_shim("dump", N, a, seed)
    assert r.returncode == 0, r.stderr + r.stdout
    r = run_ref("dump", N, b, seed)
    assert r.returncode == 0, r.stderr + r.stdout
    for name in ("summary", "TrueVariance", "Task.0"):
        assert (Path(a) / name).read_bytes() == (Path(b) / name).read_bytes(), name
    assert (Path(a) / "Task.0").stat().st_size == 56 * N ** 3
    # the unchanged reference resumes from the shim's dump, and the shim from the reference's
    r = run_ref("read", N, a, seed)
    assert r.returncode == 0, r.stderr + r.stdout
    r = run_shim("read", N, b, seed)
    assert r.returncode == 0, r.stderr + r.stdout


def test_read_dumps_rejects_a_different_run(tmp_path):
    """summary mismatch (src/fmax.c:451-478): another seed must be refused"""
    N = 16
    d = str(tmp_path / "dumps") + "/"
    assert run_shim("dump", N, d, 486604).returncode == 0
    # same files, but the reading run has another RandomSeed: host_setup(seed) differs from the summary
    text = (Path(d) / "summary").read_text().splitlines()
    text[1] = text[1].replace("486604", "12345")
    (Path(d) / "summary").write_text("\n".join(text) + "\n")
    r = run_shim("read_expect_fail", N, d, 486604)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "random seed" in r.stdout


def test_set_one_grid_geometry():
    """slab geometry of the shim's set_one_grid against what PFFT returns for a 1-D decomposition
    (src/fmax-pfft.c:95-134; real space split along x, k space along y)"""
    import ctypes
    code = r"""
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
N, task, ntasks = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
assert lib.host_setup(N, task, ntasks, 1, 1, b"/tmp/") == 0
out = (ctypes.c_long * 13)()
lib.host_geometry(out)
print(*list(out))
"""
    for N, ntasks in ((32, 1), (64, 4), (128, 8)):
        for task in (0, ntasks - 1):
            r = subprocess.run([sys.executable, "-c", code, str(rr.SHIM_HOST_LIB), str(N), str(task), str(ntasks)],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            g = [int(v) for v in r.stdout.split()]
            lx = N // ntasks
            assert g[0:3] == [lx, N, N] and g[3:6] == [task * lx, 0, 0]
            assert g[6:9] == [N, lx, N // 2 + 1] and g[9:12] == [0, task * lx, 0]
            assert g[12] == 2 * lx * N * (N // 2 + 1)
