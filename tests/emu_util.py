"""Helpers for the CPU block emulator (tests/host/emu.cpp): build, load, pitched arrays."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
SRC = HERE / "host" / "emu.cpp"
LIB = HERE / "host" / "libpinb_emu.so"

PD = ctypes.POINTER(ctypes.c_double)
PF = ctypes.POINTER(ctypes.c_float)
PI32 = ctypes.POINTER(ctypes.c_int)
PU32 = ctypes.POINTER(ctypes.c_uint)


LIB_SPLIT = HERE / "host" / "libpinb_emu_split.so"


def build_emulator(split: bool = False) -> Path:
    """split=True lowers PINB_SPLIT_ABOVE to 16, so that the 32^3 and 64^3 emulator grids take the
    decimation-in-frequency path that the product only uses for N = 2048."""
    lib = LIB_SPLIT if split else LIB
    deps = [SRC] + list((ROOT / "pinocchio_b200" / "csrc").glob("*.cuh")) + list((ROOT / "pinocchio_b200" / "csrc").glob("*.h"))
    if lib.exists() and all(lib.stat().st_mtime > d.stat().st_mtime for d in deps):
        return lib
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas"]
    if split:
        cmd += ["-DPINB_SPLIT_ABOVE=16", "-DPINB_SPLIT_ABOVE_Y=16"]
    subprocess.run(cmd + ["-o", str(lib), str(SRC)], check=True)
    return lib


def load_emulator(split: bool = False):
    lib = ctypes.CDLL(str(build_emulator(split)))
    return lib


def ptr(a, t=PD):
    return a.ctypes.data_as(t) if a is not None else None


def ptr_array(arrs, n, t=PD):
    """C array of n pointers (None -> NULL)."""
    arr = (t * n)()
    for i in range(n):
        a = arrs[i] if i < len(arrs) else None
        arr[i] = ptr(a, t) if a is not None else t()
    return arr


def twiddles(n):
    k = np.arange(n)
    return np.ascontiguousarray(np.exp(2j * np.pi * k / n).astype(np.complex128))


def pitch(N):
    return N // 2 + 8


def to_pitched_c(a):
    """[N,N,N/2+1] complex -> [N,N,P] complex (pad columns zero)."""
    N = a.shape[0]
    out = np.zeros((N, N, pitch(N)), dtype=np.complex128)
    out[:, :, : N // 2 + 1] = a
    return out


def from_pitched_c(a):
    N = a.shape[0]
    return np.ascontiguousarray(a[:, :, : N // 2 + 1])


def to_pitched_r(a):
    """[N,N,N] real -> [N,N,2P] real."""
    N = a.shape[0]
    out = np.zeros((N, N, 2 * pitch(N)), dtype=np.float64)
    out[:, :, :N] = a
    return out


def real_view(c):
    """View a pitched complex array [N,N,P] as reals [N,N,2P]; return the valid [N,N,N] part."""
    N = c.shape[0]
    return c.view(np.float64).reshape(N, N, 2 * pitch(N))[:, :, :N]


def empty_field(N):
    # garbage-filled on purpose (pad columns must never leak into results)
    return np.full((N, N, pitch(N)), 1e300 + 0j, dtype=np.complex128)


def packed_spline(lib, sp):
    """Device table of a NaturalSpline, built by the engine's own packer (csrc/spline_pack.h)."""
    lib.emu_spline_table_doubles.restype = ctypes.c_longlong
    n = sp.size
    out = np.zeros(lib.emu_spline_table_doubles(n))
    x, y = np.ascontiguousarray(sp.x), np.ascontiguousarray(sp.y)
    assert lib.emu_pack_spline(ptr(x), ptr(y), n, ptr(out)) == 0
    return out
