"""Emulated multi-rank run of the engine's schedule (pinocchio_b200/csrc/engine.cu) on the CPU.

Every rank's buffers are host arrays; the kernel bodies run under the fiber block emulator
(tests/host/emu.cpp), rank after rank, so the peer-memory scatter addressing of the x and y
passes (the fused all-to-all) is exercised without a GPU.  The barriers of the real schedule
become trivial because the emulated ranks execute each phase in turn.
"""
from __future__ import annotations

import ctypes

import numpy as np

from emu_util import PF, PI32, PU32, load_emulator, pitch, ptr, ptr_array, twiddles

HESS_JOBS = np.array([2, 0, 0, 0, 2, 1, 0, 0, 2, 1, 1, 3, 1, 0, 4, 0, 1, 5], dtype=np.int32)
HESS_KZ = np.array([0, 0, 2, 0, 1, 1], dtype=np.int32)


class EmuCluster:
    def __init__(self, N: int, nranks: int, split: bool = False):
        self.lib = load_emulator(split)   # split: long-line (N = 2048) code path on the small grids
        self.N, self.P = N, nranks
        self.lx = N // nranks
        self.Pc = pitch(N)
        self.tw = twiddles(N)
        self.norm = 1.0 / N ** 3
        R = range(nranks)
        self.kdens = [self.kfield(zero=True) for _ in R]
        self.A = [[self.rfield() for _ in R] for _ in range(3)]       # A[i][rank], R layout
        self.KV = [[self.kfield() for _ in R] for _ in range(3)]      # KV[i][rank], K layout
        self.B = [[self.rfield() for _ in R] for _ in range(6)]
        self.D = [[self.rfield() for _ in R] for _ in range(3)]
        self.Fmax = [np.zeros((self.lx, N, N), dtype=np.float32) for _ in R]
        self.Rmax = [np.zeros((self.lx, N, N), dtype=np.int32) for _ in R]
        self.sums = [np.zeros(2) for _ in R]

    # garbage-filled on purpose: pad columns / untouched regions must never leak into results
    def kfield(self, zero=False):
        a = np.zeros((self.N, self.lx, self.Pc), dtype=np.complex128)
        if not zero:
            a[...] = 1e300
        return a

    def rfield(self):
        return np.full((self.lx, self.N, self.Pc), 1e300 + 0j, dtype=np.complex128)

    # ---- assembling global views -----------------------------------------------------
    def gather_k(self, fields):
        """K-layout slabs [N][ly][P] of all ranks -> global [N][N][N/2+1]."""
        return np.concatenate([f[:, :, : self.N // 2 + 1] for f in fields], axis=1)

    def scatter_k(self, glob, fields):
        for r, f in enumerate(fields):
            f[...] = 0
            f[:, :, : self.N // 2 + 1] = glob[:, r * self.lx:(r + 1) * self.lx]

    def gather_real(self, fields):
        """R-layout real slabs -> global [N][N][N]."""
        out = [f.view(np.float64).reshape(self.lx, self.N, 2 * self.Pc)[:, :, : self.N] for f in fields]
        return np.concatenate(out, axis=0)

    # ---- kernels ----------------------------------------------------------------------
    def genic(self, seeds, pk, box):
        for r in range(self.P):
            assert self.lib.emu_genic(self.N, r, self.P, ptr(np.ascontiguousarray(seeds), PU32), ptr(pk),
                                      ctypes.c_double(box), 0, 0, ptr(self.kdens[r])) == 0

    def xpass_inv(self, src, dsts, pmask, with_nyq, gauss, scalar, green, times_i, gk=None):
        """src[rank]: K layout; dsts: {power: A-like [rank] list}; gk = (table, logkmin, dlogk, sign):
        scale-dependent growth (KFactor::gk*)."""
        flat = []
        for pw in range(3):
            flat += dsts.get(pw, [None] * self.P)
        for r in range(self.P):
            if gk is None:
                assert self.lib.emu_xpass(self.N, +1, r, self.P, ptr(src[r]), ptr_array(flat, 3 * self.P), 0, pmask,
                                          with_nyq, ptr(gauss), ctypes.c_double(scalar), green, times_i, ptr(self.tw)) == 0
            else:
                tab = np.ascontiguousarray(gk[0], dtype=np.float64)
                assert self.lib.emu_xpass_gk(self.N, +1, r, self.P, ptr(src[r]), ptr_array(flat, 3 * self.P), 0, pmask,
                                             with_nyq, ptr(gauss), ctypes.c_double(scalar), green, times_i, ptr(self.tw),
                                             ptr(tab), tab.size, ctypes.c_double(gk[1]), ctypes.c_double(gk[2]),
                                             ctypes.c_double(gk[3])) == 0

    def xpass_inv_staged(self, src, S, A, pmask, with_nyq, gauss, scalar, green, times_i):
        """The pipelined multi-GPU sweep's transposes (engine.cu run_xpass_local + transpose_dma): every rank's x pass
        writes its own x planes straight into its R-layout field A[pw][rank] and everything else into a LOCAL
        K-layout staging field S[pw][rank] (XPassParams::dst_klayout = 2); the copy engines then move block
        (source rank r -> destination rank d) = S[pw][r][d lx : (d+1) lx] into A[pw][d][:, r ly : (r+1) ly] --
        done here with NumPy slices, which is exactly what the strided cudaMemcpy2DAsync calls do."""
        lx = self.lx
        for r in range(self.P):
            flat = []
            for pw in range(3):
                flat += [S[pw][r], A[pw][r]] + [None] * (self.P - 2)
            assert self.lib.emu_xpass(self.N, +1, r, self.P, ptr(src[r]), ptr_array(flat, 3 * self.P), 2, pmask,
                                      with_nyq, ptr(gauss), ctypes.c_double(scalar), green, times_i, ptr(self.tw)) == 0
        for pw in range(3):
            if not (pmask >> pw) & 1:
                continue
            for r in range(self.P):
                for d in range(self.P):
                    if d != r:
                        A[pw][d][:, r * lx:(r + 1) * lx, :] = S[pw][r][d * lx:(d + 1) * lx, :, :]

    def ypass_inv(self, srcs, dsts, jobs, with_nyq):
        ja = np.asarray(jobs, dtype=np.int32).ravel()
        for r in range(self.P):
            assert self.lib.emu_ypass(self.N, +1, r, self.P, ptr_array([s[r] for s in srcs], 3),
                                      ptr_array([d[r] for d in dsts], 6), None, 0, ptr(ja, PI32), len(ja) // 3,
                                      with_nyq, ptr(self.tw)) == 0

    def r2c(self, src, kdst):
        """src[rank] real R layout (destroyed) -> kdst[rank] K layout."""
        for r in range(self.P):
            assert self.lib.emu_zpass_r2c(self.N, self.P, ptr(src[r]), ptr(src[r]), ptr(self.tw)) == 0
        ja = np.array([0, 0, 0], dtype=np.int32)
        for r in range(self.P):
            assert self.lib.emu_ypass(self.N, -1, r, self.P, ptr_array([src[r]], 3), None, ptr_array(kdst, self.P), 1,
                                      ptr(ja, PI32), 1, 1, ptr(self.tw)) == 0
        for r in range(self.P):
            flat = [kdst[r]] + [None] * (3 * self.P - 1)
            assert self.lib.emu_xpass(self.N, -1, r, self.P, ptr(kdst[r]), ptr_array(flat, 3 * self.P), 1, 1, 1, None,
                                      ctypes.c_double(1.0), 0, 0, ptr(self.tw)) == 0

    def hessian_collapse(self, radius, cell, spline_packed, nspl, ismooth, store_h):
        N, M = self.N, self.N // 2
        knorm = 2 * np.pi / N
        rs = radius / cell
        gauss = np.exp(-0.5 * (knorm * np.arange(M + 1)) ** 2 * rs * rs)
        dc = np.array([self.kdens[0][0, 0, 0].real * self.norm])
        self.xpass_inv(self.kdens, {0: self.A[0], 1: self.A[1], 2: self.A[2]}, 7, 0, gauss, self.norm, 1, 0)
        self.ypass_inv(self.A, self.B, HESS_JOBS, 0)
        for r in range(self.P):
            self.sums[r][:] = 0
            Br = [b[r] for b in self.B]
            hd = ptr_array(Br, 6) if store_h else None
            assert self.lib.emu_zpass_collapse(N, self.P, ptr_array(Br, 6), ptr(HESS_KZ, PI32), 0, ptr(dc),
                                               ptr(spline_packed), nspl, ismooth, ptr(self.Fmax[r], PF),
                                               ptr(self.Rmax[r], PI32), ptr(self.sums[r]), hd, ptr(self.tw)) == 0
        return sum(s[1] for s in self.sums)

    def sources(self):
        P2 = 2 * self.Pc
        for r in range(self.P):
            Hp = [b[r].view(np.float64).reshape(self.lx, self.N, P2) for b in self.B]
            S = [a[r].view(np.float64).reshape(self.lx, self.N, P2) for a in self.A]
            assert self.lib.emu_sources(self.N, self.P, ptr_array(Hp, 6), ptr(S[0]), ptr(S[1]), ptr(S[2]), 3) == 0

    def contraction(self):
        """source_3LPT_2 (A[2]) -= sum 2 w d_a d_b phi_2 * H_ab, three groups, x-dest A[0]."""
        dc = np.array([self.KV[0][0][0, 0, 0].real * self.norm])
        groups = [(2, [(0, 0, 0)], [0], [0]),
                  (1, [(0, 1, 0), (0, 0, 1)], [0, 1], [3, 4]),
                  (0, [(0, 2, 0), (0, 1, 1), (0, 0, 2)], [0, 1, 2], [1, 5, 2])]
        P2 = 2 * self.Pc
        for pw, jobs, kzp, slots in groups:
            self.xpass_inv(self.KV[0], {pw: self.A[0]}, 1 << pw, 1, None, self.norm, 1, 0)
            self.ypass_inv([self.A[0]], self.D, jobs, 1)
            kz = np.array(kzp + [0] * (6 - len(kzp)), dtype=np.int32)
            w = np.array([2.0 * (1.0 if s <= 2 else 2.0) for s in slots] + [0.0] * (6 - len(slots)))
            for r in range(self.P):
                hs = [self.B[s][r].view(np.float64).reshape(self.lx, self.N, P2) for s in slots]
                acc = self.A[2][r].view(np.float64).reshape(self.lx, self.N, P2)
                assert self.lib.emu_zpass_out(self.N, self.P, len(jobs), ptr_array([d[r] for d in self.D], 6),
                                              ptr(kz, PI32), 1, ptr(dc), 2, None, None, ptr_array(hs, 6), ptr(w),
                                              ptr(acc), ptr(self.tw)) == 0

    def displacement(self, kvec, growth, with_nyq, gk=None):
        """three float fields per rank from the K-layout k-vector kvec[rank]."""
        dc = np.array([-kvec[0][0, 0, 0].imag * self.norm])
        self.xpass_inv(kvec, {0: self.A[1], 1: self.A[0]}, 3, with_nyq, None, self.norm * growth, 1, 1, gk)
        self.ypass_inv([self.A[0], self.A[1]], self.D, [(0, 0, 0), (1, 1, 1), (1, 0, 2)], with_nyq)
        kz = np.array([0, 0, 1, 0, 0, 0], dtype=np.int32)
        out = []
        for r in range(self.P):
            V = [np.zeros((self.lx, self.N, self.N), dtype=np.float32) for _ in range(3)]
            assert self.lib.emu_zpass_out(self.N, self.P, 3, ptr_array([d[r] for d in self.D], 6), ptr(kz, PI32),
                                          with_nyq, ptr(dc), 1, None, ptr_array(V, 6, PF), None, None, None,
                                          ptr(self.tw)) == 0
            out.append(V)
        return [np.concatenate([out[r][a] for r in range(self.P)], axis=0) for a in range(3)]

    def c2r_plain(self, kvec):
        """plain c2r of K-layout kvec[rank] -> global real field (x dest A[0], y dest A[1], z in place)."""
        self.xpass_inv(kvec, {0: self.A[0]}, 1, 1, None, self.norm, 0, 0)
        self.ypass_inv([self.A[0]], [self.A[1]], [(0, 0, 0)], 1)
        kz = np.zeros(6, dtype=np.int32)
        for r in range(self.P):
            assert self.lib.emu_zpass_out(self.N, self.P, 1, ptr_array([self.A[1][r]], 6), ptr(kz, PI32), 1, None, 0,
                                          ptr_array([self.A[1][r]], 6), None, None, None, None, ptr(self.tw)) == 0
        return self.gather_real(self.A[1])
