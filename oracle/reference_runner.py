"""TEST INFRASTRUCTURE: ctypes front end of oracle/_ref/libpinocchio_ref.so, i.e. the reference's
OWN src/{fmax,fmax-pfft,LPT,collapse_times,variables}.c compiled verbatim (oracle/Makefile) and
run on one task over the PFFT/MPI/GSL stand-ins of oracle/ref_fft.c and oracle/ref_harness.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this module.  The library
keeps the reference's process-wide globals, so one process = one grid: `ReferenceRun` can be
created once; run it through `python oracle/reference_runner.py ...` (a fresh process, which
also keeps libgomp apart from torch/numpy thread pools) when several grids are needed.

    python oracle/reference_runner.py --grid 128 --steps 2 --threads 8      # prints one JSON line
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libpinocchio_ref.so"
SHIM_HOST_LIB = HERE / "_ref" / "libshim_host.so"
REFERENCE_SRC = Path("/root/reference/src")
HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
_PD = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(_PD)


def build(force: bool = False) -> Path | None:
    """make -C oracle when the reference tree is present; otherwise whatever was built before
    (the GPU box only has the prebuilt file).  Returns None when neither exists."""
    if REFERENCE_SRC.exists():
        if force and LIB.exists():
            LIB.unlink()
        # `all` also links oracle/_ref/libshim_host.so (the drop-in shim's host side) against the
        # in-tree libpinb200.so when that has been built
        target = "all" if (HERE.parent / "pinocchio_b200" / "libpinb200.so").exists() else "_ref/libpinocchio_ref.so"
        r = subprocess.run(["make", "-j", "4", "-C", str(HERE), target], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"oracle/_ref build failed:\n{r.stdout}\n{r.stderr}")
    return LIB if LIB.exists() else None


def available() -> bool:
    return LIB.exists()


class ReferenceRun:
    """One N^3 box on one task.  Files the reference writes (pinocchio.ref.FmaxPDF.out) go to
    `workdir`; its stdout chatter goes to `workdir`/log when quiet."""

    def __init__(self, N: int, box: float, radii, growth, invgrow_x, invgrow_y, threads: int | None = None,
                 variances=None, workdir: str | None = None, quiet: bool = True):
        if not LIB.exists():
            raise RuntimeError(f"{LIB} missing: run `make -C oracle` where /root/reference exists")
        self.N = N
        self.threads = threads or os.cpu_count() or 1
        self.workdir = workdir or tempfile.mkdtemp(prefix="pinref_")
        self.quiet = quiet
        self.lib = ctypes.CDLL(str(LIB))
        self.lib.ref_setup.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, _PD, _PD, _PD, ctypes.c_int]
        self.lib.ref_set_invgrow.argtypes = [ctypes.c_int, _PD, _PD]
        self.lib.ref_fetch_products.argtypes = [ctypes.c_void_p]
        self.lib.ref_fetch_kvector.argtypes = [ctypes.c_int, _PD]
        x = np.ascontiguousarray(invgrow_x, dtype=np.float64)
        y = np.ascontiguousarray(invgrow_y, dtype=np.float64)
        self.lib.ref_set_invgrow(len(x), _p(x), _p(y))
        self.radii = np.ascontiguousarray(radii, dtype=np.float64)
        var = np.ascontiguousarray(variances if variances is not None else np.ones(len(self.radii)), dtype=np.float64)
        g = np.ascontiguousarray(growth, dtype=np.float64)
        assert g.shape == (4,)
        self._in_workdir(lambda: self._check(self.lib.ref_setup(N, float(box), len(self.radii), _p(self.radii), _p(var), _p(g),
                                                                int(self.threads))))

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise RuntimeError("reference call failed")

    def _in_workdir(self, fn):
        """run fn() with cwd = workdir and (when quiet) the C-level stdout sent to workdir/log"""
        old = os.getcwd()
        os.chdir(self.workdir)
        saved = None
        try:
            if self.quiet:
                sys.stdout.flush()
                saved = os.dup(1)
                fd = os.open("log", os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o644)
                os.dup2(fd, 1)
                os.close(fd)
            return fn()
        finally:
            if saved is not None:
                ctypes.CDLL(None).fflush(None)
                os.dup2(saved, 1)
                os.close(saved)
            os.chdir(old)

    def set_kdensity(self, kd: np.ndarray):
        a = np.ascontiguousarray(kd, dtype=np.complex128)
        assert a.shape == (self.N, self.N, self.N // 2 + 1)
        self.lib.ref_set_kdensity(_p(a.view(np.float64)))

    def set_growth_tables(self, log10_growth, logkmin: float = -3.0, dlogk: float = 0.5):
        """-DSCALE_DEPENDENT: [4][NkBINS] tables for GrowingMode*(z, k) (ref_harness.c); None = off."""
        self.lib.ref_set_growth_tables.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, _PD]
        if log10_growth is None:
            self.lib.ref_set_growth_tables(0, 0.0, 1.0, None)
            return
        t = np.ascontiguousarray(log10_growth, dtype=np.float64)
        assert t.ndim == 2 and t.shape[0] == 4
        self.lib.ref_set_growth_tables(t.shape[1], float(logkmin), float(dlogk), _p(t))

    def genic(self, seed: int, pk_lattice: np.ndarray, fixed_ic: int = 0, paired_ic: int = 0) -> np.ndarray:
        """The reference's own GenIC_large (src/GenIC.c) for RandomSeed = seed; ``pk_lattice[m]`` =
        PowerSpectrum(2 pi sqrt(m)/Box), m = |n|^2 (pinocchio_b200.cosmology.pk_lattice_table).
        Fills the reference's kdensity[0] and returns a copy [N][N][N/2+1]."""
        self._pk = np.ascontiguousarray(pk_lattice, dtype=np.float64)      # must outlive the call
        self.lib.ref_genic.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _PD, ctypes.c_long]
        self._in_workdir(lambda: self._check(self.lib.ref_genic(int(seed), int(fixed_ic), int(paired_ic), _p(self._pk),
                                                                self._pk.size)))
        k = np.empty((self.N, self.N, self.N // 2 + 1), dtype=np.complex128)
        self.lib.ref_get_kdensity.argtypes = [_PD]
        self.lib.ref_get_kdensity(_p(k.view(np.float64)))
        return k

    def compute_fmax(self):
        """The reference's compute_fmax(): returns (seconds by its own cputime.fmax, TrueVariance[])."""
        sec = ctypes.c_double()
        tv = np.empty(len(self.radii))
        self._in_workdir(lambda: self._check(self.lib.ref_compute_fmax(ctypes.byref(sec), _p(tv))))
        return sec.value, tv

    def products(self, dtype) -> np.ndarray:
        assert self.lib.ref_sizeof_product() == np.dtype(dtype).itemsize
        out = np.empty(self.N ** 3, dtype=dtype)
        self.lib.ref_fetch_products(out.ctypes.data_as(ctypes.c_void_p))
        return out

    def kvector(self, which: int) -> np.ndarray:
        k = np.empty((self.N, self.N, self.N // 2 + 1), dtype=np.complex128)
        self.lib.ref_fetch_kvector(which, _p(k.view(np.float64)))
        return k

    def timers(self) -> dict:
        t = np.empty(6)
        self.lib.ref_timers(_p(t))
        return dict(zip(("fmax", "deriv", "fft", "coll", "lpt", "mem_transf"), (float(v) for v in t)))

    def fmax_pdf_file(self) -> np.ndarray:
        return np.loadtxt(Path(self.workdir) / "pinocchio.ref.FmaxPDF.out")


def synthetic_kdensity(N: int, cosmo, seed: int = 486604) -> np.ndarray:
    """The benchmark's synthetic input on the CPU side: the oracle's GenIC restatement."""
    sys.path.insert(0, str(HERE.parent))
    from oracle import pinocchio_oracle as po
    return po.genic(N, N / 0.7, seed, cosmo.PowerSpectrum)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    sys.path.insert(0, str(HERE.parent))
    from pinocchio_b200.cosmology import Cosmology
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    N = args.grid
    g = (cosmo.GrowingMode(0.0), cosmo.GrowingMode_2LPT(0.0), cosmo.GrowingMode_3LPT_1(0.0), cosmo.GrowingMode_3LPT_2(0.0))
    run = ReferenceRun(N, N / 0.7, HMF_RADII, g, cosmo.sp_invgrow.x, cosmo.sp_invgrow.y, threads=args.threads)
    run.set_kdensity(synthetic_kdensity(N, cosmo))
    for _ in range(args.warmup):
        run.compute_fmax()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        run.compute_fmax()
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    print(json.dumps({"grid": N, "threads": args.threads, "steps": args.steps, "seconds_per_step": t,
                      "mcells_per_s": N ** 3 / t / 1e6, "timers_last_step": run.timers()}))


if __name__ == "__main__":
    main()
