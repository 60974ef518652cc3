/* pinb200 -- C ABI of the B200-native collapse-time engine for PINOCCHIO V5.1.
 *
 * Drop-in boundary: the reference has no plugin system; the boundary is
 * the set of C symbols the rest of PINOCCHIO calls on this path.  shim/fmax_b200.c defines
 * those reference symbols (GenIC_large, compute_fmax, compute_displacements, ...) on top of the
 * entry points below; INTEGRATION.md shows the binding.  Plain pointers and sizes only, no
 * CUDA/torch types.  Every function returns 0 on success and non-zero on failure, like the
 * reference (src/pinocchio.c:229-230,259-263); pinb200_last_error() gives the message.
 *
 * There is no CPU fallback: every entry point that computes runs sm_100a kernels and fails
 * if no CUDA device is usable.
 */
#ifndef PINB200_H
#define PINB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pinb200_ctx pinb200_ctx;

#define PINB200_NBINS 210 /* NBINS, reference src/pinocchio.h:65 */

/* Run descriptor: the reference globals this path reads (SURVEY.md 8b "inputs read from
 * globals"), passed as runtime values instead of -D switches. */
typedef struct {
  int grid_size;      /* params.GridSize[0]; cubic, power of two in [32, 2048]              */
  double box_size;    /* MyGrids[0].BoxSize = params.BoxSize_htrue, true Mpc                */
  int random_seed;    /* params.RandomSeed                                                  */
  int fixed_ic;       /* params.FixedIC                                                     */
  int paired_ic;      /* params.PairedIC                                                    */
  int lpt_order;      /* 1: Zel'dovich only, 2: -DTWO_LPT, 3: -DTWO_LPT -DTHREE_LPT          */
  int rank, nranks;   /* ThisTask, NTasks: slab decomposition along x (src/initialization.c:1317-1325) */
  int device;         /* CUDA device ordinal for this rank                                  */
} pinb200_desc;

/* Layout of one product_data record (reference src/pinocchio.h:233-259): byte offsets of the
 * members inside the record, -1 for members compiled out; prodfloat_bytes = sizeof(PRODFLOAT). */
typedef struct {
  size_t stride;      /* sizeof(product_data)                                               */
  int prodfloat_bytes;/* 4 (default) or 8 (-DDOUBLE_PRECISION_PRODUCTS)                      */
  int off_Rmax, off_Fmax, off_Vel, off_Vel_2LPT, off_Vel_3LPT_1, off_Vel_3LPT_2;
} pinb200_product_layout;

/* cputime_data members this path fills (reference src/pinocchio.h:368-378), in seconds of
 * device time measured with CUDA events. */
typedef struct {
  double dens, fmax, deriv, fft, coll, lpt, mem_transf;
  double per_radius[64];   /* wall of each smoothing radius (src/fmax.c:140-146)             */
  double hess_x, hess_y, hess_z; /* device seconds summed over radii: x pass (fused k-space kernel),
                                    y pass, z pass + collapse                                  */
  double disp_sources, disp_vel; /* displacement stage: sources + k-vectors; 4 x first derivatives */
  unsigned long long kernel_launches;  /* kernels launched by this context so far            */
  double sort_ms;                      /* last pinb200_collapsed_cells: selection + sort on the device, milliseconds */
  double disp_x;                       /* displacement stage: the four inverse x passes of the first derivatives, seconds
                                          (with pinb200_displacements_scaledep: xpass_growthk_kernel, growth rate per mode) */
  double xfer;                         /* multi-GPU sweep: copy-engine transposes (first barrier passed -> all local copies
                                          issued by this rank done), seconds, summed over the radii; they run under the
                                          collapse pass of the previous radius */
} pinb200_timers;

/* ---- life cycle ---------------------------------------------------------------------- */
int pinb200_create(const pinb200_desc* desc, pinb200_ctx** out);
int pinb200_destroy(pinb200_ctx* ctx);
const char* pinb200_last_error(const pinb200_ctx* ctx); /* ctx may be NULL: creation errors */
/* Multi-GPU (nranks > 1): one process per GPU of one box.  The distributed-FFT transposes that
 * PFFT does with MPI all-to-alls (src/fmax-pfft.c:197,211) are stores into the peers' memory
 * over NVLink; each rank exports one cudaIpc handle and maps its peers' handles.  The caller
 * moves the 64-byte handles between processes (MPI_Allgather in the reference's world,
 * torch.distributed.all_gather in bench.py): all_handles = nranks * 64 bytes, ordered by rank.
 * Every rank must then make the same sequence of compute calls (they contain barriers). */
#define PINB200_IPC_HANDLE_BYTES 64
int pinb200_ipc_handle(pinb200_ctx* ctx, void* handle64);
int pinb200_connect(pinb200_ctx* ctx, const void* all_handles);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*). */
int pinb200_set_stream(pinb200_ctx* ctx, void* cuda_stream);
int pinb200_synchronize(pinb200_ctx* ctx);

/* ---- tables computed by the unchanged host code ----------------------------------------- */
/* P(k) on the integer lattice: pk[m] = PowerSpectrum(2*pi*sqrt(m)/box_size), m = 0..(N/2)^2
 * (replaces the per-mode PowerSpectrum() call of src/GenIC.c:283). */
int pinb200_set_power_table(pinb200_ctx* ctx, const double* pk, size_t n);
/* Smoothing ladder: Smoothing.Nsmooth, Smoothing.Radius[] in true Mpc (src/initialization.c:386-435). */
int pinb200_set_smoothing(pinb200_ctx* ctx, int nsmooth, const double* radius);
/* Knots of the inverse growing mode spline, x = log10 D, y = log10 a: SPLINE[SP_INVGROW]
 * (src/cosmo.c:401) when ismooth < 0, else SPLINE_INVGROW[ismooth] (src/initialization.c:1704-1708).
 * The natural-cubic-spline coefficients are recomputed here as gsl_interp_cspline does. */
int pinb200_set_invgrow_spline(pinb200_ctx* ctx, int ismooth, const double* x, const double* y, int n);

/* ---- tabulated collapse times: -DTABULATED_CT, with ELL_CLASSIC or ELL_SNG ---- */
/* The reference fills, for every smoothing radius, a table F(delta, x, y) of nbins_d * nbins_xy^2
 * points with ell() -- ell_classic, or the numerical ellipsoidal collapse ell_sng: one 9-variable
 * rkf45 integration per point (initialize_collapse_times, src/collapse_times.c:824-1046; ell_sng
 * :315-400; sng_system :239-290) -- and evaluates per cell four cubic splines in delta blended
 * bilinearly in (x, y) (interpolate_collapse_time, BILINEAR_SPLINE, :1132-1222).  model is the type
 * code of the CTtable file header (write_CTtable_header, :1307-1326): 1 = ELL_CLASSIC tabulated,
 * 3 = ELL_SNG standard gravity, 4 = ELL_SNG with the Hu-Sawicki f(R) force modification (MOD_GRAV_FR,
 * ForceModification, :294-311). */
#define PINB200_CT_CLASSIC 1
#define PINB200_CT_SNG 3
#define PINB200_CT_SNG_FR 4
typedef struct {
  int model;
  int nbins_d, nbins_xy;      /* CT_NBINS_D (100, at most 128), CT_NBINS_XY (50) */
  double range_x;             /* CT_RANGE_X (3.5): bin_x = range_x / nbins_xy */
  const double* delta_vector; /* nbins_d increasing knots; NULL = pinb200_ct_delta_vector() */
  /* ELL_SNG only: OmegaMatter(z), OmegaLambda(z) of src/cosmo.c:1675-1718 for a cosmological constant:
   * E^2(z) = omega_rad (1+z)^4 + omega0 (1+z)^3 + omega_k (1+z)^2 + omega_lambda */
  double omega0, omega_lambda, omega_rad, omega_k;
  /* model 4 only: FR0, H_over_c = 100 / SPEEDOFLIGHT (src/cosmo.c:109) and, per smoothing radius, the size handed
   * to sng_system: Smoothing.Radius[ismooth], the previous radius for the last one (src/collapse_times.c:362-372) */
  double fr0, h_over_c;
  const double* fr_size;
} pinb200_ct_desc;
/* delta_vector of the reference's compiled sampling (CT_EXPO 1.75, CT_SQUEEZE 1.2, CT_RANGE_D 7,
 * CT_DELTA0 -1; src/collapse_times.c:781-787, 836-877); host code, no device needed */
int pinb200_ct_delta_vector(double* delta_vector, int nbins_d);
/* Switch pinb200_fmax / pinb200_collapse_cells to tabulated collapse times (desc = NULL: back to the
 * direct ell_classic evaluation).  variance[ismooth] = Smoothing.Variance (the table is sampled in
 * units of its square root); d_in[ismooth] = GrowingMode(1/1e-5 - 1, k(R_ismooth)), ELL_SNG only
 * (src/collapse_times.c:345-353).  tables = nsmooth * nbins_d * nbins_xy^2 doubles in the reference's
 * CT_table order (delta fastest, then x, then y) as read from a CTtableFile, or NULL: the tables are
 * computed on the device -- with ELL_CLASSIC from the inverse-growth splines set before this call.
 * Needs pinb200_set_smoothing first. */
int pinb200_set_collapse_tables(pinb200_ctx* ctx, const pinb200_ct_desc* desc, const double* variance, const double* d_in,
                                const double* tables);
/* CT_table of one radius (what the reference writes to pinocchio.<run>.CTtable.out) */
int pinb200_download_collapse_table(pinb200_ctx* ctx, int ismooth, double* table);

/* ---- the hot path ------------------------------------------------------------------------ */
/* Seed plane given by the caller instead of the spiral of gsl_rng_mt19937(RandomSeed) outputs
 * (src/GenIC.c:840-855,953-973): seeds[j * GridSize + i] = seed of the (kx = i, ky = j) column, the
 * whole plane on every rank, as SEEDTABLE of src/GenIC.c:229-235.  The shim uses it for the
 * `MimicOldSeed` parameter-file option (internal.mimic_original_seedtable: the N-GenIC table of
 * src/GenIC.c:493-537, copied column for column by copy_seeds_subregion, :990-1012).  Call before
 * pinb200_genic; n must be GridSize^2. */
int pinb200_set_seed_plane(pinb200_ctx* ctx, const unsigned int* seeds, size_t n);
/* GenIC_large (src/GenIC.c:73-460): fills kdensity on the device. */
int pinb200_genic(pinb200_ctx* ctx);
/* Alternative to genic: supply / fetch kdensity as [x (N)][y_local (N/nranks)][N/2+1] complex128.
 * On one rank this is the reference's host layout (src/GenIC.c:384); on several ranks k-space is
 * split along y (PFFT's transposed layout, src/fmax-pfft.c:265-281). */
int pinb200_upload_kdensity(pinb200_ctx* ctx, const double* kdensity);
int pinb200_download_kdensity(pinb200_ctx* ctx, double* kdensity);
/* compute_fmax (src/fmax.c:36-190) without the final displacement call: loop over the
 * smoothing radii; true_variance[ismooth] = Smoothing.TrueVariance (may be NULL). */
int pinb200_fmax(pinb200_ctx* ctx, double* true_variance);
/* compute_displacements(compute_sources, 0, z) (src/fmax.c:292-367): growth[] = growth_rate of
 * src/fmax-pfft.c:344-364 for ScaleDep.order 1..4 at the segment redshift, i.e.
 * {GrowingMode, GrowingMode_2LPT, GrowingMode_3LPT_1 (negative), GrowingMode_3LPT_2}. */
int pinb200_displacements(pinb200_ctx* ctx, int compute_sources, const double growth[4]);
/* The same for -DSCALE_DEPENDENT builds, where growth_rate depends on |k| (src/fmax-pfft.c:340-364,
 * InterpolateGrowth src/cosmo.c:1728-1757).  log10_growth[o*nk + j], o = ScaleDep.order-1 = 0..3,
 * j = 0..nk-1, is my_spline_eval(SPLINE[SP_GROW1|SP_GROW2|SP_GROW31|SP_GROW32 + j], -log10(1+z))
 * at the segment redshift z (evaluated by the unchanged host cosmology); nk = NkBINS,
 * logkmin = LOGKMIN, dlogk = DELTALOGK (src/def_splines.h:40-42).  The device interpolates
 * linearly in log10|k| (|k| in grid units, as the reference passes it), clamps below kmin and
 * above kmax, takes 10^ and applies the minus sign of GrowingMode_3LPT_1 (src/cosmo.c:1810).
 * The sources (compute_sources = 1) are computed with growth 1 as in the reference
 * (ScaleDep.order = 0, src/fmax.c:308-309, src/LPT.c:53-54). */
int pinb200_displacements_scaledep(pinb200_ctx* ctx, int compute_sources, int nk, double logkmin, double dlogk,
                                   const double* log10_growth);
/* Fmax_PDF (src/fmax.c:509-550): local histogram, counts[PINB200_NBINS]. */
int pinb200_fmax_pdf(pinb200_ctx* ctx, unsigned long long* counts);

/* ---- results ----------------------------------------------------------------------------- */
/* Pack cells [cell_begin, cell_begin+ncells) of the local slab (index = z + N*(y + N*x_local),
 * src/pinocchio.h:84-85) into host AoS records.  Only the members this path computes are written
 * (Rmax, Fmax, Vel* -- as compute_collapse_times and write_from_rvector_to_products do,
 * src/collapse_times.c:587-590, src/fmax-pfft.c:563-631): other bytes of the records (the *_prev
 * members of -DRECOMPUTE_DISPLACEMENTS, zacc/group_ID of -DSNAPSHOT) keep the caller's values. */
int pinb200_download_products(pinb200_ctx* ctx, void* products, const pinb200_product_layout* layout,
                              size_t cell_begin, size_t ncells);
/* Hand-off to the fragmentation: the local cells with Fmax >= f_last -- the
 * ones distribute() keeps (src/distribute.c:58-175,547-600) -- as cell indices
 * (z + N*(y + N*x_local)) in order of descending Fmax, the order sort_and_organize gives frag[]
 * (src/fragment.c:484-520); equal Fmax in ascending cell index.  *count = number of such cells;
 * at most `capacity` indices are written (cell_index_out may be NULL to query the count). */
int pinb200_collapsed_cells(pinb200_ctx* ctx, float f_last, unsigned int* cell_index_out, size_t capacity, size_t* count);
/* The same selection and order, started under the displacement stage: Fmax and Rmax are final when pinb200_fmax
 * returns, the displacement fields only after pinb200_displacements.  _begin counts the selected cells (*count, one
 * synchronisation) and then lets the compaction, the sort and the download of the index list (at most `capacity`
 * entries into cell_index_out) run on a side stream, and the download of the whole Fmax field into fmax_out
 * (local cells floats; may be NULL) on the copy engine; it returns at once.  The caller then runs
 * pinb200_displacements, and pinb200_handoff_end waits for the list.  Host arrays should be pinned (a pageable
 * destination makes the copies synchronous).  pinb200_collapsed_cells = _begin + _end. */
int pinb200_handoff_begin(pinb200_ctx* ctx, float f_last, float* fmax_out, unsigned int* cell_index_out, size_t capacity, size_t* count);
int pinb200_handoff_end(pinb200_ctx* ctx);
/* frag[first .. first+n) as sort_and_organize leaves it: the product_data records of the cells
 * listed by the last pinb200_collapsed_cells call (which must have been given an output array), in
 * that order, gathered on the device -- only collapsed cells cross PCIe and the host does not sort. */
int pinb200_download_products_sorted(pinb200_ctx* ctx, void* products, const pinb200_product_layout* layout,
                                     size_t first, size_t n);
/* Structure-of-arrays access for tests: which = 0 Fmax(f32) 1 Rmax(i32) 2..4 Vel 5..7 Vel_2LPT
 * 8..10 Vel_3LPT_1 11..13 Vel_3LPT_2; dst holds N^3 (local) 4-byte values. */
int pinb200_download_field(pinb200_ctx* ctx, int which, void* dst);
int pinb200_get_timers(pinb200_ctx* ctx, pinb200_timers* t);

/* On-disk formats whose payload is the device SoA, written to an open file descriptor
 * straight from the device through two pinned staging buffers (copy engine and write(2) overlap), without a
 * host products[] array:
 *  - pinb200_write_products: the AoS records of cells [cell_begin, +ncells) as DumpProducts/Task.<rank> holds them
 *    (src/fmax.c:418-420: fwrite(products, sizeof(product_data), total_local_size)); members of the record that
 *    this path does not compute are written as zero bytes;
 *  - pinb200_write_block: the payload of one block of the timeless snapshot for those cells, as initialize_FMAX /
 *    _RMAX / _ZEL / _2LPT / _3LPT_1 / _3LPT_2 fill it (src/write_snapshot.c:695-860): float, int, or three
 *    interleaved floats (AuxStruct) per cell.  Block headers and the INFO block stay with the caller. */
#define PINB200_BLOCK_FMAX 0
#define PINB200_BLOCK_RMAX 1
#define PINB200_BLOCK_ZEL 2
#define PINB200_BLOCK_2LPT 3
#define PINB200_BLOCK_3LPT_1 4
#define PINB200_BLOCK_3LPT_2 5
int pinb200_write_products(pinb200_ctx* ctx, int fd, const pinb200_product_layout* layout, size_t cell_begin, size_t ncells);
int pinb200_write_block(pinb200_ctx* ctx, int fd, int block, size_t cell_begin, size_t ncells);

/* ---- finer-grained entry points (reference function granularity; used by the parity tests) */
/* forward_transform / reverse_transform (src/fmax-pfft.c:191-228) on host arrays:
 * real [N/nranks (local x)][N][N] doubles <-> half-complex [N][N/nranks (local y)][N/2+1];
 * reverse includes the 1/N^3.  Collective when nranks > 1 (every rank passes its slab). */
int pinb200_fft_r2c(pinb200_ctx* ctx, const double* real_in, double* cplx_out);
int pinb200_fft_c2r(pinb200_ctx* ctx, const double* cplx_in, double* real_out);
/* compute_second_derivatives(R) (src/fmax.c:225-258) of the resident kdensity: six real
 * [N/nranks][N][N] double fields in slot order xx,yy,zz,xy,xz,yz written to hessian_out
 * (6 * local cells; may be NULL).  With R = 0 the fields also stay resident as the input of the
 * LPT sources: this is recompute_sd = 1 of compute_displacements (src/fmax.c:301-319, special
 * mode 3 of src/pinocchio.c:186), to be followed by pinb200_displacements(compute_sources = 1). */
int pinb200_second_derivatives(pinb200_ctx* ctx, double radius, double* hessian_out);
/* inverse_collapse_time over ncells Hessians given as six arrays (SoA), ismooth selects the
 * spline; F_out[ncells] doubles (src/collapse_times.c:679-776). */
int pinb200_collapse_cells(pinb200_ctx* ctx, int ismooth, const double* hessian6, size_t ncells, double* F_out);
/* kvector_2LPT / kvector_3LPT_1 / kvector_3LPT_2 (src/LPT.c:98-172) in the host layout of
 * kdensity; which = 0,1,2.  Valid after pinb200_displacements(compute_sources=1). */
int pinb200_download_kvector(pinb200_ctx* ctx, int which, double* kvec);

/* ---- start-up integrals of the scale-dependent growth (next row f4) -------------------------------------------
 * Replaces the three gsl_integration_qags loops of set_scaledep_GM (src/initialization.c:1594-1601, :1742-1748,
 * :1886-1892): for every smoothing radius r, every time knot i of SPLINE[SP_TIME] and the three quantities
 * q = 0 density (Gaussian window, IntegrandForSDDensVariance :1439), 1 displacement (top-hat,
 * IntegrandForSDDisplVariance :1449), 2 velocity (top-hat, times fomega^2, IntegrandForSDVelVariance :1489)
 *     out[(q * nsmooth + r) * ntimes + i] = sqrt( Int dlog10k  P(k) D(t_i,k)^2 [f(t_i,k)^2] W(k R_r)^2 k^p / (2 pi^2) )
 * i.e. the reference's vector[i] before its normalisation, all 3 x nsmooth x ntimes of them in two kernel launches.
 * The quadrature is the caller's: nodes in log10 k over the reference's interval [-4, nyquist] and, per node,
 * a = weight * PowerSpectrum(k) * k^p / (2 pi^2) from the host cosmology (p = 3 density, 1 displacement/velocity);
 * shim/scaledep_gm_b200.c uses composite 8-point Gauss-Legendre.  Growth enters as the k-bin tables InterpolateGrowth
 * (src/cosmo.c:1728-1755) interpolates: log10 GrowingMode and fomega at the NkBINS bins and the time knots.
 * Bisection for k_GM_*, normalisation and the SPLINE_INVGROW set-up stay with the caller.  No context needed. */
typedef struct {
  int device;                 /* CUDA device ordinal                                                   */
  int nnodes;                 /* quadrature nodes                                                      */
  const double* logk;         /* [nnodes] log10 k                                                      */
  const double* a_dens;       /* [nnodes] weight * P(k) * k^3 / (2 pi^2)                               */
  const double* a_disp;       /* [nnodes] weight * P(k) * k / (2 pi^2)                                 */
  int nkbins, ntimes;         /* NkBINS (1 without -DSCALE_DEPENDENT), NBINS                           */
  double logkmin, dlogk;      /* LOGKMIN, DELTALOGK (src/def_splines.h:41-42)                          */
  const double* log10_growth; /* [nkbins][ntimes] SPLINE[SP_GROW1 + kk] at the time knots              */
  const double* fomega;       /* [nkbins][ntimes] SPLINE[SP_FOMEGA1 + kk] there                        */
  int nsmooth;                /* Smoothing.Nsmooth, <= 64                                              */
  const double* radius_dens;  /* [nsmooth] Smoothing.Radius (Gaussian)                                 */
  const double* radius_disp;  /* [nsmooth] Smoothing.Rad_GM (top-hat)                                  */
} pinb200_sdgm_desc;
int pinb200_scaledep_variances(const pinb200_sdgm_desc* desc, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PINB200_H */
