"""Host-side mirror of the reference interface for the collapse-time path, over the C ABI.

``Pinocchio`` keeps the reference's function names and argument meaning for this path
(``GenIC_large``, ``compute_fmax``, ``compute_displacements``, ``Fmax_PDF``, ... --
src/pinocchio.h:544-566,575-577,637-648) and its error convention (0 = OK, non-zero = failure,
message available).  All compute happens in libpinb200.so (sm_100a kernels); this module only
marshals tables and buffers through ``ctypes``.  There is no CPU fallback: loading fails loudly
if the library is missing, and every call fails if no B200 is visible.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .cosmology import Cosmology, SmoothingLadder, pk_lattice_table, set_smoothing

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libpinb200.so"
NBINS = 210

_PD = ctypes.POINTER(ctypes.c_double)


class Desc(ctypes.Structure):
    _fields_ = [("grid_size", ctypes.c_int), ("box_size", ctypes.c_double), ("random_seed", ctypes.c_int),
                ("fixed_ic", ctypes.c_int), ("paired_ic", ctypes.c_int), ("lpt_order", ctypes.c_int),
                ("rank", ctypes.c_int), ("nranks", ctypes.c_int), ("device", ctypes.c_int)]


class ProductLayout(ctypes.Structure):
    _fields_ = [("stride", ctypes.c_size_t), ("prodfloat_bytes", ctypes.c_int), ("off_Rmax", ctypes.c_int),
                ("off_Fmax", ctypes.c_int), ("off_Vel", ctypes.c_int), ("off_Vel_2LPT", ctypes.c_int),
                ("off_Vel_3LPT_1", ctypes.c_int), ("off_Vel_3LPT_2", ctypes.c_int)]


class CTDesc(ctypes.Structure):
    """pinb200_ct_desc: tabulated collapse times (-DTABULATED_CT, src/collapse_times.c:780-1346)"""
    _fields_ = [("model", ctypes.c_int), ("nbins_d", ctypes.c_int), ("nbins_xy", ctypes.c_int),
                ("range_x", ctypes.c_double), ("delta_vector", ctypes.POINTER(ctypes.c_double)),
                ("omega0", ctypes.c_double), ("omega_lambda", ctypes.c_double), ("omega_rad", ctypes.c_double),
                ("omega_k", ctypes.c_double), ("fr0", ctypes.c_double), ("h_over_c", ctypes.c_double),
                ("fr_size", ctypes.POINTER(ctypes.c_double))]


CT_CLASSIC, CT_SNG, CT_SNG_FR = 1, 3, 4   # type codes of the CTtable file header (src/collapse_times.c:1307-1326)
H_OVER_C = 100.0 / 299792.458             # H_over_c = 100 / SPEEDOFLIGHT (src/cosmo.c:109)
CT_NBINS_D, CT_NBINS_XY, CT_RANGE_X = 100, 50, 3.5   # src/collapse_times.c:781-787


class SdgmDesc(ctypes.Structure):
    """pinb200_sdgm_desc (include/pinb200.h): inputs of the batched set_scaledep_GM integrals."""
    _fields_ = [("device", ctypes.c_int), ("nnodes", ctypes.c_int), ("logk", ctypes.POINTER(ctypes.c_double)),
                ("a_dens", ctypes.POINTER(ctypes.c_double)), ("a_disp", ctypes.POINTER(ctypes.c_double)),
                ("nkbins", ctypes.c_int), ("ntimes", ctypes.c_int), ("logkmin", ctypes.c_double), ("dlogk", ctypes.c_double),
                ("log10_growth", ctypes.POINTER(ctypes.c_double)), ("fomega", ctypes.POINTER(ctypes.c_double)),
                ("nsmooth", ctypes.c_int), ("radius_dens", ctypes.POINTER(ctypes.c_double)),
                ("radius_disp", ctypes.POINTER(ctypes.c_double))]


class Timers(ctypes.Structure):
    _fields_ = [("dens", ctypes.c_double), ("fmax", ctypes.c_double), ("deriv", ctypes.c_double),
                ("fft", ctypes.c_double), ("coll", ctypes.c_double), ("lpt", ctypes.c_double),
                ("mem_transf", ctypes.c_double), ("per_radius", ctypes.c_double * 64),
                ("hess_x", ctypes.c_double), ("hess_y", ctypes.c_double), ("hess_z", ctypes.c_double),
                ("disp_sources", ctypes.c_double), ("disp_vel", ctypes.c_double),
                ("kernel_launches", ctypes.c_ulonglong), ("sort_ms", ctypes.c_double),
                ("disp_x", ctypes.c_double), ("xfer", ctypes.c_double)]


# every symbol include/pinb200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "pinb200_create", "pinb200_destroy", "pinb200_last_error", "pinb200_set_stream", "pinb200_synchronize",
    "pinb200_ipc_handle", "pinb200_connect",
    "pinb200_set_power_table", "pinb200_set_smoothing", "pinb200_set_invgrow_spline", "pinb200_set_seed_plane", "pinb200_genic",
    "pinb200_upload_kdensity", "pinb200_download_kdensity", "pinb200_fmax", "pinb200_displacements",
    "pinb200_displacements_scaledep", "pinb200_collapsed_cells", "pinb200_download_products_sorted",
    "pinb200_fmax_pdf", "pinb200_download_products", "pinb200_download_field", "pinb200_get_timers",
    "pinb200_write_products", "pinb200_write_block", "pinb200_handoff_begin", "pinb200_handoff_end",
    "pinb200_fft_r2c", "pinb200_fft_c2r", "pinb200_second_derivatives", "pinb200_collapse_cells",
    "pinb200_download_kvector",
    "pinb200_ct_delta_vector", "pinb200_set_collapse_tables", "pinb200_download_collapse_table",
    "pinb200_scaledep_variances",
]

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libpinb200.so (built in-tree by ``pinocchio_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    # PINB200_LIB: a build of the same sources with other compile-time options (tests only)
    path = Path(os.environ["PINB200_LIB"]) if os.environ.get("PINB200_LIB") else LIB_PATH
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run `python -m pinocchio_b200.build` "
                           "(there is no CPU fallback for the collapse-time path)")
    lib = ctypes.CDLL(str(path))
    lib.pinb200_last_error.restype = ctypes.c_char_p
    lib.pinb200_last_error.argtypes = [ctypes.c_void_p]
    lib.pinb200_create.argtypes = [ctypes.POINTER(Desc), ctypes.POINTER(ctypes.c_void_p)]
    lib.pinb200_destroy.argtypes = [ctypes.c_void_p]
    lib.pinb200_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.pinb200_synchronize.argtypes = [ctypes.c_void_p]
    lib.pinb200_ipc_handle.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.pinb200_connect.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.pinb200_set_power_table.argtypes = [ctypes.c_void_p, _PD, ctypes.c_size_t]
    lib.pinb200_set_smoothing.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD]
    lib.pinb200_set_invgrow_spline.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD, _PD, ctypes.c_int]
    lib.pinb200_set_seed_plane.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint), ctypes.c_size_t]
    lib.pinb200_genic.argtypes = [ctypes.c_void_p]
    lib.pinb200_upload_kdensity.argtypes = [ctypes.c_void_p, _PD]
    lib.pinb200_download_kdensity.argtypes = [ctypes.c_void_p, _PD]
    lib.pinb200_fmax.argtypes = [ctypes.c_void_p, _PD]
    lib.pinb200_displacements.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD]
    lib.pinb200_displacements_scaledep.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                   ctypes.c_double, _PD]
    lib.pinb200_collapsed_cells.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.POINTER(ctypes.c_uint), ctypes.c_size_t,
                                            ctypes.POINTER(ctypes.c_size_t)]
    lib.pinb200_handoff_begin.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint),
                                          ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.pinb200_handoff_end.argtypes = [ctypes.c_void_p]
    lib.pinb200_write_products.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ProductLayout), ctypes.c_size_t, ctypes.c_size_t]
    lib.pinb200_write_block.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t]
    lib.pinb200_download_products_sorted.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ProductLayout),
                                                     ctypes.c_size_t, ctypes.c_size_t]
    lib.pinb200_fmax_pdf.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.pinb200_download_products.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ProductLayout),
                                              ctypes.c_size_t, ctypes.c_size_t]
    lib.pinb200_download_field.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.pinb200_get_timers.argtypes = [ctypes.c_void_p, ctypes.POINTER(Timers)]
    lib.pinb200_fft_r2c.argtypes = [ctypes.c_void_p, _PD, _PD]
    lib.pinb200_fft_c2r.argtypes = [ctypes.c_void_p, _PD, _PD]
    lib.pinb200_second_derivatives.argtypes = [ctypes.c_void_p, ctypes.c_double, _PD]
    lib.pinb200_collapse_cells.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD, ctypes.c_size_t, _PD]
    lib.pinb200_download_kvector.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD]
    lib.pinb200_ct_delta_vector.argtypes = [_PD, ctypes.c_int]
    lib.pinb200_set_collapse_tables.argtypes = [ctypes.c_void_p, ctypes.POINTER(CTDesc), _PD, _PD, _PD]
    lib.pinb200_download_collapse_table.argtypes = [ctypes.c_void_p, ctypes.c_int, _PD]
    lib.pinb200_scaledep_variances.argtypes = [ctypes.POINTER(SdgmDesc), _PD]
    _lib = lib
    return lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_PD)


def gauss_legendre_nodes(lo: float, hi: float, npanels: int = 512, order: int = 8, breaks=()):
    """Composite Gauss-Legendre nodes and weights on [lo, hi]: the quadrature shim/scaledep_gm_b200.c hands to
    pinb200_scaledep_variances (order 8, about 512 panels of the reference's interval [-4, nyquist] in log10 k).
    `breaks`: points where the integrand has kinks (the k bins of InterpolateGrowth's piecewise-linear blend): panel
    edges are put there, every segment getting its share of the panels (at least one)."""
    x, w = np.polynomial.legendre.leggauss(order)
    cuts = [lo] + sorted(b for b in breaks if lo < b < hi) + [hi]
    edges = [np.array([lo])]
    for a, b in zip(cuts[:-1], cuts[1:]):
        m = max(1, int(np.floor(npanels * (b - a) / (hi - lo) + 0.5)))
        edges.append(a + (b - a) * np.arange(1, m + 1) / m)
    edges = np.concatenate(edges)
    edges[-1] = hi
    h = 0.5 * (edges[1:] - edges[:-1])
    c = 0.5 * (edges[1:] + edges[:-1])
    return (c[:, None] + h[:, None] * x[None, :]).ravel(), (h[:, None] * w[None, :]).ravel()


def scaledep_variances(logk, a_dens, a_disp, log10_growth, fomega, logkmin, dlogk, radius_dens, radius_disp,
                       device: int = 0) -> np.ndarray:
    """set_scaledep_GM's integrals (src/initialization.c:1594-1601, :1742-1748, :1886-1892), all of them in one device
    call: returns out[3][nsmooth][ntimes] = sqrt of the density / displacement / velocity integrals, the reference's
    `vector` before its normalisation.  logk: quadrature nodes in log10 k; a_dens, a_disp: weight * P(k) * k^3 (k) /
    (2 pi^2) per node from the host cosmology; log10_growth, fomega: [nkbins][ntimes] k-bin tables at the time knots."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (logk, a_dens, a_disp, log10_growth, fomega, radius_dens, radius_disp)]
    lk, ad, ap, lg, fo, rd, rp = arrs
    if lg.ndim != 2 or fo.shape != lg.shape or ad.shape != lk.shape or ap.shape != lk.shape or rd.shape != rp.shape:
        raise ValueError("scaledep_variances: inconsistent shapes")
    d = SdgmDesc(device, lk.size, _dp(lk), _dp(ad), _dp(ap), lg.shape[0], lg.shape[1], float(logkmin), float(dlogk),
                 _dp(lg), _dp(fo), rd.size, _dp(rd), _dp(rp))
    out = np.zeros((3, rd.size, lg.shape[1]))
    lib = load_library()
    if lib.pinb200_scaledep_variances(ctypes.byref(d), _dp(out)):
        raise PinocchioError((lib.pinb200_last_error(None) or b"pinb200_scaledep_variances failed").decode())
    return out


# product_data for -DTWO_LPT -DTHREE_LPT, float products (src/pinocchio.h:233-259): 56 bytes
PRODUCT_DTYPE_3LPT = np.dtype([("Rmax", "<i4"), ("Fmax", "<f4"), ("Vel", "<f4", 3), ("Vel_2LPT", "<f4", 3),
                               ("Vel_3LPT_1", "<f4", 3), ("Vel_3LPT_2", "<f4", 3)])
# the same with -DDOUBLE_PRECISION_PRODUCTS (PRODFLOAT = double, src/pinocchio.h:225-231): Fmax at offset 8, 112 bytes
PRODUCT_DTYPE_3LPT_DOUBLE = np.dtype({"names": ["Rmax", "Fmax", "Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"],
                                      "formats": ["<i4", "<f8", ("<f8", 3), ("<f8", 3), ("<f8", 3), ("<f8", 3)],
                                      "offsets": [0, 8, 16, 40, 64, 88], "itemsize": 112})
FIELD_INDEX = {"Fmax": 0, "Rmax": 1, "Vel": 2, "Vel_2LPT": 5, "Vel_3LPT_1": 8, "Vel_3LPT_2": 11}


class PinocchioError(RuntimeError):
    pass


@dataclass
class RunConfig:
    """The reference globals this path reads (params, MyGrids[0], Smoothing, outputs)."""
    GridSize: int = 128
    BoxSize_htrue: float = 128.0 / 0.7      # true Mpc (params.BoxSize_htrue, src/initialization.c:235-245)
    RandomSeed: int = 486604
    FixedIC: int = 0
    PairedIC: int = 0
    lpt_order: int = 3                      # TWO_LPT + THREE_LPT
    zlast: float = 0.0                      # outputs.zlast
    segment_redshift: float = 0.0           # ScaleDep.z[0]


class Pinocchio:
    """One rank of the collapse-time path.  Method names follow the reference."""

    def __init__(self, cfg: RunConfig, cosmo: Cosmology | None = None, device: int = 0,
                 smoothing: SmoothingLadder | None = None, rank: int = 0, nranks: int = 1, group=None):
        """``group``: a torch.distributed process group (or None for the default) used to move the
        cudaIpc handles and to sum the per-rank reductions when nranks > 1 -- the role MPI plays in
        the reference (MPI_Reduce / MPI_Bcast, src/collapse_times.c:656-667, src/fmax.c:527)."""
        self.lib = load_library()
        self.rank, self.nranks, self.group = rank, nranks, group
        self.cfg = cfg
        self.cosmo = cosmo if cosmo is not None else Cosmology()
        self.N = int(cfg.GridSize)
        d = Desc(self.N, float(cfg.BoxSize_htrue), int(cfg.RandomSeed), int(cfg.FixedIC), int(cfg.PairedIC),
                 int(cfg.lpt_order), rank, nranks, device)
        h = ctypes.c_void_p()
        if self.lib.pinb200_create(ctypes.byref(d), ctypes.byref(h)):
            raise PinocchioError(self.lib.pinb200_last_error(None).decode())
        self.h = h
        self.lx = self.N // nranks          # local x extent (real space) == local y extent (k space)
        if nranks > 1:
            from .distributed import connect_peers
            connect_peers(self)
        self.CellSize = cfg.BoxSize_htrue / self.N
        self.Smoothing = smoothing if smoothing is not None else set_smoothing(self.cosmo, self.CellSize, cfg.zlast)
        self.TrueVariance = np.zeros(self.Smoothing.Nsmooth)
        self._set_tables()

    # -- plumbing -----------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise PinocchioError(self.lib.pinb200_last_error(self.h).decode())

    def _set_tables(self):
        pk = np.ascontiguousarray(pk_lattice_table(self.cosmo, self.N, self.cfg.BoxSize_htrue))
        self._ck(self.lib.pinb200_set_power_table(self.h, _dp(pk), pk.size))
        r = np.ascontiguousarray(self.Smoothing.Radius, dtype=np.float64)
        self._ck(self.lib.pinb200_set_smoothing(self.h, r.size, _dp(r)))
        sp = self.cosmo.sp_invgrow
        x, y = np.ascontiguousarray(sp.x), np.ascontiguousarray(sp.y)
        self._ck(self.lib.pinb200_set_invgrow_spline(self.h, -1, _dp(x), _dp(y), x.size))

    def close(self):
        if getattr(self, "h", None):
            self.lib.pinb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.pinb200_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.pinb200_synchronize(self.h))

    # -- reference entry points ---------------------------------------------------------------
    def set_seed_plane(self, seeds: np.ndarray) -> None:
        """SEEDTABLE[j * N + i] of src/GenIC.c:229-235 given by the caller (the `MimicOldSeed` table of
        src/GenIC.c:493-537 in the shim); the default is the spiral table of generate_seeds_plane."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32).ravel()
        self._ck(self.lib.pinb200_set_seed_plane(self.h, seeds.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), seeds.size))

    def _host_barrier(self) -> None:
        """Every rank enters a collective device call together (MPI_Barrier in the shim): the cross-GPU
        barrier inside the library orders memory, it is not meant to absorb seconds of host-side skew."""
        if self.nranks > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def GenIC_large(self, ThisGrid: int = 0) -> int:
        """src/GenIC.c:73-460."""
        self._ck(self.lib.pinb200_genic(self.h))
        return 0

    def compute_fmax(self, displacements: bool = True) -> int:
        """src/fmax.c:36-190: radii loop, then compute_displacements(1, 0, ScaleDep.z[0])."""
        tv = np.zeros(self.Smoothing.Nsmooth)
        self._host_barrier()
        self._ck(self.lib.pinb200_fmax(self.h, _dp(tv)))
        if self.nranks > 1:
            from .distributed import allreduce_sum
            tv = allreduce_sum(tv, self.group)
        self.TrueVariance = tv
        if displacements:
            self.compute_displacements(1, 0, self.cfg.segment_redshift)
        return 0

    def growth_rates(self, redshift: float) -> np.ndarray:
        c = self.cosmo
        return np.array([c.GrowingMode(redshift), c.GrowingMode_2LPT(redshift), c.GrowingMode_3LPT_1(redshift),
                         c.GrowingMode_3LPT_2(redshift)])

    def set_scale_dependent_growth(self, log10_growth_of_z, logkmin: float = -3.0, dlogk: float = 0.5):
        """-DSCALE_DEPENDENT (src/def_splines.h:40-42): ``log10_growth_of_z(z)`` returns the [4][NkBINS]
        values my_spline_eval(SPLINE[SP_GROW1|2|31|32 + j], -log10(1+z)) that InterpolateGrowth
        (src/cosmo.c:1728-1757) interpolates in log10 k; None switches back to scale-independent growth."""
        self._scaledep = None if log10_growth_of_z is None else (log10_growth_of_z, float(logkmin), float(dlogk))

    def compute_displacements(self, compute_sources: int, recompute_sd: int, redshift: float) -> int:
        """src/fmax.c:292-367.  recompute_sd: the R = 0 second derivatives are computed first
        (special mode 3, src/pinocchio.c:186: displacements without an Fmax sweep)."""
        self._host_barrier()
        if recompute_sd:
            self._ck(self.lib.pinb200_second_derivatives(self.h, 0.0, None))
        sd = getattr(self, "_scaledep", None)
        if sd is not None:
            tab = np.ascontiguousarray(sd[0](redshift), dtype=np.float64)
            if tab.ndim != 2 or tab.shape[0] != 4:
                raise PinocchioError("scale-dependent growth tables must have shape [4][NkBINS]")
            self._ck(self.lib.pinb200_displacements_scaledep(self.h, int(compute_sources), tab.shape[1], sd[1], sd[2],
                                                             _dp(tab)))
            return 0
        g = self.growth_rates(redshift)
        self._ck(self.lib.pinb200_displacements(self.h, int(compute_sources), _dp(g)))
        return 0

    # -- -DTABULATED_CT -------------------------------------------------------------------------
    def initialize_collapse_times(self, model: int = CT_CLASSIC, tables: np.ndarray | None = None, nbins_d: int = CT_NBINS_D,
                                  nbins_xy: int = CT_NBINS_XY, delta_vector: np.ndarray | None = None, fr0: float = 0.0,
                                  d_in: np.ndarray | None = None) -> int:
        """initialize_collapse_times for every smoothing radius (src/collapse_times.c:824-1046): from now on
        compute_fmax / inverse_collapse_time interpolate F in per-radius tables -- computed on the device
        with ell_classic (model CT_CLASSIC) or with the ELL_SNG ellipsoid integration (CT_SNG), or taken
        from ``tables`` [Nsmooth][nbins_xy][nbins_xy][nbins_d] (a CTtableFile).  model None: back to the
        direct evaluation.  CT_SNG_FR: Hu-Sawicki f(R) force modification with ``fr0`` = FR0 (-DMOD_GRAV_FR); ``d_in``
        then carries the scale-dependent GrowingMode(1/1e-5 - 1, k(R)) of the host cosmology per radius."""
        if model is None:
            self._ck(self.lib.pinb200_set_collapse_tables(self.h, None, None, None, None))
            return 0
        c = self.cosmo
        d = CTDesc(int(model), int(nbins_d), int(nbins_xy), CT_RANGE_X, None, c.p.Omega0, c.p.OmegaLambda, c.OmegaRad, c.OmegaK,
                   float(fr0), H_OVER_C, None)
        # the size handed to sng_system: the radius, the previous one for the last (R = 0) (src/collapse_times.c:362-372)
        size = np.ascontiguousarray(self.Smoothing.Radius, dtype=np.float64).copy()
        if size.size > 1:
            size[-1] = size[-2]
        d.fr_size = _dp(size)
        dv = None
        if delta_vector is not None:
            dv = np.ascontiguousarray(delta_vector, dtype=np.float64)
            assert dv.size == nbins_d
            d.delta_vector = _dp(dv)
        var = np.ascontiguousarray(self.Smoothing.Variance, dtype=np.float64)
        # GrowingMode(1/amin - 1, .) of ell_sng (src/collapse_times.c:345-353); scale-independent growth
        d_in = np.full(var.size, c.GrowingMode(1.0 / 1.0e-5 - 1.0)) if d_in is None else np.ascontiguousarray(d_in, dtype=np.float64)
        assert d_in.size == var.size
        tab = None
        if tables is not None:
            tab = np.ascontiguousarray(tables, dtype=np.float64)
            assert tab.size == var.size * nbins_d * nbins_xy * nbins_xy
        self._ck(self.lib.pinb200_set_collapse_tables(self.h, ctypes.byref(d), _dp(var), _dp(d_in),
                                                      _dp(tab) if tab is not None else None))
        self._ct_shape = (nbins_xy, nbins_xy, nbins_d)
        return 0

    def collapse_table(self, ismooth: int) -> np.ndarray:
        """CT_table of one radius, [iy][ix][id] (index id + nd*(ix + nxy*iy), src/collapse_times.c:966-977)."""
        out = np.zeros(self._ct_shape)
        self._ck(self.lib.pinb200_download_collapse_table(self.h, int(ismooth), _dp(out)))
        return out

    @staticmethod
    def ct_delta_vector(nbins_d: int = CT_NBINS_D) -> np.ndarray:
        dv = np.zeros(nbins_d)
        if load_library().pinb200_ct_delta_vector(_dp(dv), nbins_d):
            raise PinocchioError("pinb200_ct_delta_vector failed")
        return dv

    def Fmax_PDF(self) -> np.ndarray:
        """src/fmax.c:509-550; returns the 210 counts."""
        c = (ctypes.c_ulonglong * NBINS)()
        self._ck(self.lib.pinb200_fmax_pdf(self.h, c))
        pdf = np.array(list(c), dtype=np.int64)
        if self.nranks > 1:
            from .distributed import allreduce_sum
            pdf = allreduce_sum(pdf, self.group)
        return pdf.astype(np.uint64)

    def collapsed_cells(self, Flast: float) -> np.ndarray:
        """Local cells with Fmax >= Flast as indices z + N*(y + N*x_local), in order of descending Fmax
        (ties: ascending index): the selection of src/distribute.c and the order of
        sort_and_organize (src/fragment.c:484-520), computed on the device."""
        n = ctypes.c_size_t(0)
        self._ck(self.lib.pinb200_collapsed_cells(self.h, float(Flast), None, 0, ctypes.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            self._ck(self.lib.pinb200_collapsed_cells(self.h, float(Flast), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)),
                                                      out.size, ctypes.byref(n)))
        self._sorted_n = int(n.value)
        return out

    def sorted_products(self, first: int = 0, n: int | None = None, dtype=PRODUCT_DTYPE_3LPT) -> np.ndarray:
        """frag[] of src/fragment.c:484-520: records of the cells of the last collapsed_cells() call, in
        that order (descending Fmax), gathered on the device."""
        if n is None:
            n = self._sorted_n - first
        out = np.zeros(n, dtype=dtype)
        f = dtype.fields
        off = lambda name: f[name][1] if name in f else -1
        lay = ProductLayout(dtype.itemsize, dtype["Fmax"].itemsize, off("Rmax"), off("Fmax"), off("Vel"), off("Vel_2LPT"),
                            off("Vel_3LPT_1"), off("Vel_3LPT_2"))
        if n:
            self._ck(self.lib.pinb200_download_products_sorted(self.h, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lay),
                                                               first, n))
        return out

    # -- data movement ----------------------------------------------------------------------
    def write_kdensity(self, kdensity: np.ndarray):
        """Upload kdensity[0]: [x][y_local][N/2+1] complex128 (the reference layout on one rank)."""
        a = np.ascontiguousarray(kdensity, dtype=np.complex128)
        assert a.shape == (self.N, self.lx, self.N // 2 + 1)
        self._ck(self.lib.pinb200_upload_kdensity(self.h, a.view(np.float64).ctypes.data_as(_PD)))

    def read_kdensity(self) -> np.ndarray:
        a = np.zeros((self.N, self.lx, self.N // 2 + 1), dtype=np.complex128)
        self._ck(self.lib.pinb200_download_kdensity(self.h, a.view(np.float64).ctypes.data_as(_PD)))
        return a

    def read_kvector(self, which: int) -> np.ndarray:
        a = np.zeros((self.N, self.lx, self.N // 2 + 1), dtype=np.complex128)
        self._ck(self.lib.pinb200_download_kvector(self.h, which, a.view(np.float64).ctypes.data_as(_PD)))
        return a

    def field(self, name: str, comp: int = 0) -> np.ndarray:
        idx = FIELD_INDEX[name] + (comp if name.startswith("Vel") else 0)
        dt = np.int32 if name == "Rmax" else np.float32
        a = np.zeros((self.lx, self.N, self.N), dtype=dt)       # local slab: x in [rank*lx, (rank+1)*lx)
        self._ck(self.lib.pinb200_download_field(self.h, idx, a.ctypes.data_as(ctypes.c_void_p)))
        return a

    def products(self, cell_begin: int = 0, ncells: int | None = None, dtype=PRODUCT_DTYPE_3LPT) -> np.ndarray:
        """products[] records in the reference AoS layout (src/pinocchio.h:233-263)."""
        if ncells is None:
            ncells = self.lx * self.N ** 2 - cell_begin
        out = np.zeros(ncells, dtype=dtype)
        f = dtype.fields
        off = lambda n: f[n][1] if n in f else -1
        lay = ProductLayout(dtype.itemsize, dtype["Fmax"].itemsize, off("Rmax"), off("Fmax"), off("Vel"), off("Vel_2LPT"),
                            off("Vel_3LPT_1"), off("Vel_3LPT_2"))
        self._ck(self.lib.pinb200_download_products(self.h, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lay),
                                                    cell_begin, ncells))
        return out

    BLOCKS = {"FMAX": 0, "RMAX": 1, "ZEL ": 2, "2LPT": 3, "31PT": 4, "32PT": 5}      # block names of src/write_snapshot.c

    def write_products(self, fileno: int, cell_begin: int = 0, ncells: int | None = None, dtype=PRODUCT_DTYPE_3LPT) -> None:
        """DumpProducts/Task.<rank> payload (src/fmax.c:418-420) written to an open descriptor from the device SoA."""
        if ncells is None:
            ncells = self.lx * self.N ** 2 - cell_begin
        f = dtype.fields
        off = lambda n: f[n][1] if n in f else -1
        lay = ProductLayout(dtype.itemsize, dtype["Fmax"].itemsize, off("Rmax"), off("Fmax"), off("Vel"), off("Vel_2LPT"),
                            off("Vel_3LPT_1"), off("Vel_3LPT_2"))
        self._ck(self.lib.pinb200_write_products(self.h, int(fileno), ctypes.byref(lay), cell_begin, ncells))

    def write_block(self, fileno: int, name: str, cell_begin: int = 0, ncells: int | None = None) -> None:
        """Payload of one timeless-snapshot block (initialize_FMAX ... initialize_3LPT_2, src/write_snapshot.c:695-860)."""
        if ncells is None:
            ncells = self.lx * self.N ** 2 - cell_begin
        self._ck(self.lib.pinb200_write_block(self.h, int(fileno), self.BLOCKS[name], cell_begin, ncells))

    def timers(self) -> Timers:
        t = Timers()
        self._ck(self.lib.pinb200_get_timers(self.h, ctypes.byref(t)))
        return t

    # -- finer-grained reference functions (parity tests) ----------------------------------
    def forward_transform(self, r: np.ndarray) -> np.ndarray:
        """src/fmax-pfft.c:191-200 on this rank's slab: real [lx][N][N] -> half-complex [N][ly][N/2+1]."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        assert r.shape == (self.lx, self.N, self.N)
        out = np.zeros((self.N, self.lx, self.N // 2 + 1), dtype=np.complex128)
        self._ck(self.lib.pinb200_fft_r2c(self.h, _dp(r), out.view(np.float64).ctypes.data_as(_PD)))
        return out

    def reverse_transform(self, c: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=np.complex128)
        assert c.shape == (self.N, self.lx, self.N // 2 + 1)
        out = np.zeros((self.lx, self.N, self.N), dtype=np.float64)
        self._ck(self.lib.pinb200_fft_c2r(self.h, c.view(np.float64).ctypes.data_as(_PD), _dp(out)))
        return out

    def compute_second_derivatives(self, R: float) -> np.ndarray:
        out = np.zeros((6, self.lx, self.N, self.N), dtype=np.float64)
        self._host_barrier()
        self._ck(self.lib.pinb200_second_derivatives(self.h, float(R), _dp(out)))
        return out

    def inverse_collapse_time(self, hessian6: np.ndarray, ismooth: int = 0) -> np.ndarray:
        h = np.ascontiguousarray(hessian6, dtype=np.float64)
        assert h.shape[0] == 6
        n = h[0].size
        F = np.zeros(n)
        self._ck(self.lib.pinb200_collapse_cells(self.h, ismooth, _dp(h), n, _dp(F)))
        return F.reshape(h.shape[1:])
