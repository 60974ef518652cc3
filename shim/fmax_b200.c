/* fmax_b200.c -- drop-in replacement of PINOCCHIO's collapse-time path on top of libpinb200.
 *
 * Build PINOCCHIO V5.1 with this file INSTEAD OF src/fmax.c, src/fmax-pfft.c, src/collapse_times.c,
 * src/LPT.c and src/GenIC.c (see INTEGRATION.md for the Makefile lines) and link -lpinb200.
 * It defines the reference's own external symbols for this path (src/pinocchio.h:544-566,
 * 575-577, 637-648) by marshalling the reference globals into the C ABI of include/pinb200.h.
 * Everything else -- parameter file, cosmology tables, memory arena, products[] layout,
 * fragmentation, output -- is the unchanged reference code.
 *
 * This file is host glue in the reference's own language (C); it contains no numerics.
 */
#include "pinocchio.h"
#include "def_splines.h"
#include "pinb200.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

static pinb200_ctx *pinb = NULL;
#ifdef TABULATED_CT
static int shim_collapse_tables(int onlycompute);
#endif

static int pinb_fail(const char *where)
{
  printf("ERROR on task %d: %s: %s\n", ThisTask, where, pinb200_last_error(pinb));
  fflush(stdout);
  return 1;
}

/* product_data as compiled (whatever -D switches are on): offsets taken from the real struct */
static pinb200_product_layout product_layout(void)
{
  pinb200_product_layout L;
  L.stride = sizeof(product_data);
  L.prodfloat_bytes = (int)sizeof(PRODFLOAT);
  L.off_Rmax = (int)offsetof(product_data, Rmax);
  L.off_Fmax = (int)offsetof(product_data, Fmax);
  L.off_Vel = (int)offsetof(product_data, Vel);
  L.off_Vel_2LPT = L.off_Vel_3LPT_1 = L.off_Vel_3LPT_2 = -1;
#ifdef TWO_LPT
  L.off_Vel_2LPT = (int)offsetof(product_data, Vel_2LPT);
#ifdef THREE_LPT
  L.off_Vel_3LPT_1 = (int)offsetof(product_data, Vel_3LPT_1);
  L.off_Vel_3LPT_2 = (int)offsetof(product_data, Vel_3LPT_2);
#endif
#endif
  return L;
}

/* ---- hand-off to the fragmentation ------------------------------------------------------------
 * What the unchanged CPU stage reads of products[] (src/distribute.c:547-600, 685-698): Fmax of EVERY cell (the
 * distribution map keeps the cells with Fmax >= outputs.Flast) and the whole record of the cells it keeps.  So
 * only those cross PCIe: the Fmax field (4 bytes per cell) and, selected and ordered on the device
 * (pinb200_collapsed_cells), the records of the collapsed cells -- about 60 % of the cells of a z = 0 run, 44 of
 * the 60 GB of a 1024^3 box -- which are scattered into their places in products[] here.  Used when a record
 * holds nothing but members of this path (no SNAPSHOT / RECOMPUTE_DISPLACEMENTS members, whose consumers read
 * every cell); PINB200_FULL_PRODUCTS=1 forces the plain copy of every record.  (DumpProducts is written from the
 * device, see dump_products below, so it does not need the full host array either.) */
static int use_compact_handoff(void)
{
#if defined(SNAPSHOT) || defined(RECOMPUTE_DISPLACEMENTS)
  return 0;
#else
  const char *env = getenv("PINB200_FULL_PRODUCTS");
  if (env && atoi(env))
    return 0;
  return outputs.Flast > 0.0;
#endif
}

static int download_products_compact(const pinb200_product_layout *L)
{
  const size_t n = MyGrids[0].total_local_size, chunk = (size_t)1 << 22;
  size_t count = 0, first, k;
  float *fm, flast;
  unsigned int *idx;
  unsigned char *rec;

  /* (float)Fmax >= (double)Flast  <=>  Fmax >= the smallest float that is not below Flast */
  flast = (float)outputs.Flast;
  if ((double)flast < (double)outputs.Flast)
    flast = nextafterf(flast, INFINITY);

  fm = (float *)malloc(n * sizeof(float));
  if (!fm)
    return 1;
  if (pinb200_download_field(pinb, 0, fm))
  {
    free(fm);
    return pinb_fail("download Fmax");
  }
#pragma omp parallel for
  for (k = 0; k < n; k++)
    products[k].Fmax = fm[k];
  free(fm);

  if (pinb200_collapsed_cells(pinb, flast, NULL, 0, &count))
    return pinb_fail("collapsed cells");
  if (!count)
    return 0;
  idx = (unsigned int *)malloc(count * sizeof(unsigned int));
  rec = (unsigned char *)malloc((count < chunk ? count : chunk) * L->stride);
  if (!idx || !rec)
    return 1;
  if (pinb200_collapsed_cells(pinb, flast, idx, count, &count))
    return pinb_fail("collapsed cells");
  for (first = 0; first < count; first += chunk)
  {
    const size_t m = count - first < chunk ? count - first : chunk;
    if (pinb200_download_products_sorted(pinb, rec, L, first, m))
      return pinb_fail("download collapsed records");
#pragma omp parallel for
    for (k = 0; k < m; k++)
      memcpy(&products[idx[first + k]], rec + k * L->stride, L->stride);
  }
  free(rec);
  free(idx);
  if (!ThisTask)
    printf("[%s] products[]: Fmax of %zu cells and the records of the %zu with Fmax >= %g (compact hand-off)\n", fdate(), n, count,
           (double)flast);
  return 0;
}

/* ---- replaces src/fmax-pfft.c:80-134 ---------------------------------------------------- */
int set_one_grid(int ThisGrid)
{
  /* same slab geometry PFFT returns for NTasks <= GridSize (src/initialization.c:1317-1325) */
  grid_data *G = &MyGrids[ThisGrid];
  ptrdiff_t N = G->GSglobal[_x_];
  if (N % NTasks)
  {
    if (!ThisTask)
      printf("ERROR: GridSize must be divisible by the number of tasks for the GPU path\n");
    return 1;
  }
  G->norm = 1.0 / (double)G->Ntotal;
  G->CellSize = G->BoxSize / N;
  G->GSlocal[_x_] = N / NTasks; G->GSlocal[_y_] = N; G->GSlocal[_z_] = N;
  G->GSstart[_x_] = ThisTask * (N / NTasks); G->GSstart[_y_] = 0; G->GSstart[_z_] = 0;
  G->GSlocal_k[_x_] = N; G->GSlocal_k[_y_] = N / NTasks; G->GSlocal_k[_z_] = N / 2 + 1;   /* y slabs in k space */
  G->GSstart_k[_x_] = 0; G->GSstart_k[_y_] = ThisTask * (N / NTasks); G->GSstart_k[_z_] = 0;
  G->total_local_size = (unsigned int)(G->GSlocal[_x_] * N * N);
  G->total_local_size_fft = (unsigned int)(2 * G->GSlocal[_x_] * N * (N / 2 + 1));
  G->off = 0;
  return 0;
}

/* ---- replaces src/fmax-pfft.c:139-188: creates the device context and uploads the tables -- */
int compute_fft_plans(void)
{
  pinb200_desc d;
  int N = params.GridSize[0], i, ngpu = 8;
  char *env = getenv("PINB200_GPUS_PER_NODE");
  if (env) ngpu = atoi(env);

  d.grid_size = N;
  d.box_size = MyGrids[0].BoxSize;            /* params.BoxSize_htrue, src/initialization.c:490 */
  d.random_seed = params.RandomSeed;
  d.fixed_ic = params.FixedIC;
  d.paired_ic = params.PairedIC;
  d.lpt_order = 1;
#ifdef TWO_LPT
  d.lpt_order = 2;
#ifdef THREE_LPT
  d.lpt_order = 3;
#endif
#endif
  d.rank = ThisTask;
  d.nranks = NTasks;
  d.device = ThisTask % ngpu;
  if (pinb200_create(&d, &pinb))
    return pinb_fail("pinb200_create");

  if (NTasks > 1)
  {
    /* the only inter-process traffic of the GPU path on the host: one 64-byte handle per rank */
    unsigned char mine[PINB200_IPC_HANDLE_BYTES];
    unsigned char *all = (unsigned char *)malloc((size_t)NTasks * PINB200_IPC_HANDLE_BYTES);
    if (pinb200_ipc_handle(pinb, mine))
      return pinb_fail("pinb200_ipc_handle");
    MPI_Allgather(mine, PINB200_IPC_HANDLE_BYTES, MPI_BYTE, all, PINB200_IPC_HANDLE_BYTES, MPI_BYTE, MPI_COMM_WORLD);
    if (pinb200_connect(pinb, all))
      return pinb_fail("pinb200_connect");
    free(all);
    MPI_Barrier(MPI_COMM_WORLD);
  }

  /* P(k) on the integer lattice |n|^2 (replaces the per-mode call of src/GenIC.c:283) */
  {
    size_t m, n = (size_t)(N / 2) * (N / 2) + 1;
    double *pk = (double *)malloc(n * sizeof(double));
    pk[0] = 0.0;
    for (m = 1; m < n; m++)
      pk[m] = PowerSpectrum(2. * PI * sqrt((double)m) / MyGrids[0].BoxSize);
    if (pinb200_set_power_table(pinb, pk, n))
      return pinb_fail("pinb200_set_power_table");
    free(pk);
  }

  if (pinb200_set_smoothing(pinb, Smoothing.Nsmooth, Smoothing.Radius))
    return pinb_fail("pinb200_set_smoothing");

  /* inverse growing mode: knots only, the cspline coefficients are recomputed by the library */
  if (pinb200_set_invgrow_spline(pinb, -1, SPLINE[SP_INVGROW]->x, SPLINE[SP_INVGROW]->y, (int)SPLINE[SP_INVGROW]->size))
    return pinb_fail("pinb200_set_invgrow_spline");
#if defined(SCALE_DEPENDENT) && defined(ELL_CLASSIC)
  for (i = 0; i < Smoothing.Nsmooth; i++)
    if (pinb200_set_invgrow_spline(pinb, i, SPLINE_INVGROW[i]->x, SPLINE_INVGROW[i]->y, (int)SPLINE_INVGROW[i]->size))
      return pinb_fail("pinb200_set_invgrow_spline");
#else
  (void)i;
#endif
  return 0;
}

int finalize_fft(void)
{
#ifndef RECOMPUTE_DISPLACEMENTS
  /* with RECOMPUTE_DISPLACEMENTS the k-vectors must survive until the last segment
     (src/allocations.c:578-600): the context is then released at exit.  With DumpProducts it lives until
     dump_products() has written Task.<rank> from the device */
  if (params.DumpProducts && pinb)
    return 0;
  pinb200_destroy(pinb);
  pinb = NULL;
#endif
  return 0;
}

/* ---- replaces src/GenIC.c:73-460 ------------------------------------------------------------ */
/* `MimicOldSeed` (internal.mimic_original_seedtable, src/ReadParamfile.c:253-255): the seed of column
 * (i, j) comes from the N-GenIC table of src/GenIC.c:493-537 instead of the spiral -- N/2 square rings
 * grown inwards from the four corners of the plane, ring r of a corner being r cells of a column and
 * then r + 1 cells of a row, each cell 0x7fffffff * gsl_rng_uniform(random_generator) with the host's
 * own generator (ranlxd1, src/initialization.c:73) seeded with RandomSeed.  copy_seeds_subregion
 * (src/GenIC.c:990-1012) hands the table over column for column, so it IS the seed plane. */
static int mimic_old_seed_plane(void)
{
  const size_t n = (size_t)params.GridSize[0];
  unsigned int *t = (unsigned int *)calloc(n * n, sizeof(unsigned int));
  int rc;
  if (t == NULL)
    return 1;
  gsl_rng_set(random_generator, params.RandomSeed);
  for (size_t r = 0; r < n / 2; r++)
    for (int corner = 0; corner < 4; corner++)
    {
      const int flip_col = corner & 1, flip_row = corner >> 1; /* (0,0), (N,0), (0,N), (N,N) */
      const size_t ring_c = flip_col ? n - 1 - r : r, ring_r = flip_row ? n - 1 - r : r;
      for (size_t m = 0; m < r; m++) /* the column piece of the ring */
        t[(flip_row ? n - 1 - m : m) * n + ring_c] = 0x7fffffff * gsl_rng_uniform(random_generator);
      for (size_t m = 0; m < r + 1; m++) /* the row piece, corner cell included */
        t[ring_r * n + (flip_col ? n - 1 - m : m)] = 0x7fffffff * gsl_rng_uniform(random_generator);
    }
  rc = pinb200_set_seed_plane(pinb, t, n * n);
  free(t);
  return rc;
}

int GenIC_large(int ThisGrid)
{
  (void)ThisGrid;
  if ((internal.dump_seedplane || internal.dump_kdensity) && !ThisTask)
    printf("[B200 path] DumpSeedPlane / DumpKDensity are debugging dumps of the replaced src/GenIC.c:152,451 "
           "and src/fmax-pfft.c:669-707: not written\n");
  if (internal.mimic_original_seedtable && mimic_old_seed_plane())
    return pinb_fail("GenIC_large (MimicOldSeed seed plane)");
  if (pinb200_genic(pinb))
    return pinb_fail("GenIC_large");
  return 0;
}

/* ---- replaces src/fmax.c:292-367 ------------------------------------------------------------- */
int compute_displacements(int compute_sources, int recompute_sd, double redshift)
{
  double growth[4], t0 = MPI_Wtime();
  pinb200_product_layout L = product_layout();

  if (recompute_sd)
  {
    /* second derivatives at R = 0 are not in place (no Fmax sweep before): src/fmax.c:301-319 */
    double t1 = MPI_Wtime();
    ScaleDep.order = 0;
    ScaleDep.redshift = 0.0;
    MPI_Barrier(MPI_COMM_WORLD); /* collective on the device: absorb the ranks' host-side skew here */
    if (pinb200_second_derivatives(pinb, 0.0, NULL))
      return pinb_fail("compute_second_derivatives");
    cputime.deriv += MPI_Wtime() - t1;
  }
  /* Every rank enters the device collective together: with RECOMPUTE_DISPLACEMENTS this function is
     re-entered from fragment.c right after the load-imbalanced build_groups/distribute, where a skew of
     seconds is normal; the cross-GPU barrier inside the library is not meant to wait that long */
  MPI_Barrier(MPI_COMM_WORLD);
#ifdef SCALE_DEPENDENT
  (void)growth;
  {
    /* growth_rate depends on |k| (src/fmax-pfft.c:340-364): hand InterpolateGrowth's k-bin splines,
       evaluated at this redshift, to the device (src/cosmo.c:1728-1757; SP_GROW* src/def_splines.h:55-58) */
    static const int first[4] = {SP_GROW1, SP_GROW2, SP_GROW31, SP_GROW32};
    double tab[4 * NkBINS];
    int o, j;
    for (o = 0; o < 4; o++)
      for (j = 0; j < NkBINS; j++)
        tab[o * NkBINS + j] = my_spline_eval(SPLINE[first[o] + j], -log10(1. + redshift), ACCEL[first[o] + j]);
    if (pinb200_displacements_scaledep(pinb, compute_sources, NkBINS, LOGKMIN, DELTALOGK, tab))
      return pinb_fail("compute_displacements");
  }
#else
  /* growth_rate of src/fmax-pfft.c:344-364 for ScaleDep.order = 1..4 */
  growth[0] = GrowingMode(redshift, params.k_for_GM);
  growth[1] = GrowingMode_2LPT(redshift, params.k_for_GM);
  growth[2] = GrowingMode_3LPT_1(redshift, params.k_for_GM);
  growth[3] = GrowingMode_3LPT_2(redshift, params.k_for_GM);
  if (pinb200_displacements(pinb, compute_sources, growth))
    return pinb_fail("compute_displacements");
#endif
  cputime.lpt += MPI_Wtime() - t0;

  /* products[] is carved from the reference's arena; it is only filled here, never retained */
  t0 = MPI_Wtime();
  if (use_compact_handoff())
  {
    if (download_products_compact(&L))
      return 1;
  }
  else if (pinb200_download_products(pinb, products, &L, 0, MyGrids[0].total_local_size))
    return pinb_fail("download products");
  cputime.mem_transf += MPI_Wtime() - t0;
  return 0;
}

/* ---- replaces src/fmax.c:509-550 -------------------------------------------------------------- */
int Fmax_PDF(void)
{
  unsigned long long my_counter[NBINS], counter[NBINS], coll = 0;
  int i;
  if (pinb200_fmax_pdf(pinb, my_counter))
    return pinb_fail("Fmax_PDF");
  MPI_Reduce(my_counter, counter, NBINS, MPI_UNSIGNED_LONG_LONG, MPI_SUM, 0, MPI_COMM_WORLD);
  if (!ThisTask)
  {
    char filename[LBLENGTH];
    FILE *file;
    for (i = 10; i < NBINS; i++)
      coll += counter[i];
    printf("[%s] Number of collapsed particles to z=0: %Lu\n", fdate(), coll);
    sprintf(filename, "pinocchio.%s.FmaxPDF.out", params.RunFlag);
    file = fopen(filename, "w");
    fprintf(file, "# Fmax PDF over %Lu particles\n", MyGrids[0].Ntotal);
    fprintf(file, "# 1-2: F interval\n# 3: number of particles in that interval\n#\n");
    for (i = 0; i < NBINS; i++)
      fprintf(file, " %6.1f   %6.1f  %Lu\n", (double)i / 10., (i == NBINS - 1 ? 999.0 : (double)(i + 1) / 10.), counter[i]);
    fclose(file);
  }
  return 0;
}

/* ---- replaces src/fmax.c:36-190 ----------------------------------------------------------------- */
int compute_fmax(void)
{
  int ismooth;
  double *tv = (double *)malloc(Smoothing.Nsmooth * sizeof(double));
  pinb200_timers tm;

  cputime.fmax = MPI_Wtime();
  if (!ThisTask)
    printf("[%s] First part: computation of collapse times (B200 path)\n", fdate());

  ScaleDep.order = 0;
  ScaleDep.redshift = 0.0;
#ifdef TABULATED_CT
  {
    /* initialize_collapse_times for every radius, src/fmax.c:102-118 */
    double cputmp = MPI_Wtime();
    if (shim_collapse_tables(0))
      return 1;
    cputmp = MPI_Wtime() - cputmp;
    if (!ThisTask)
    {
      if (strcmp(params.CTtableFile, "none"))
        printf("[%s] Collapse times read from file %s\n", fdate(), params.CTtableFile);
      else
        printf("[%s] Collapse times computed for interpolation, cpu time =%f s\n", fdate(), cputmp);
    }
  }
#endif
  MPI_Barrier(MPI_COMM_WORLD); /* (rank 0 alone may just have written the CTtable file) */
  if (pinb200_fmax(pinb, tv))
    return pinb_fail("compute_fmax");
  /* the library returns this rank's share of Sum(delta^2)/Ntotal (src/collapse_times.c:656-670) */
  MPI_Allreduce(tv, Smoothing.TrueVariance, Smoothing.Nsmooth, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  free(tv);

  pinb200_get_timers(pinb, &tm);
  if (!ThisTask)
    for (ismooth = 0; ismooth < Smoothing.Nsmooth; ismooth++)
      printf("[%s] Completed, R=%6.3f, expected sigma: %7.4f, computed sigma: %7.4f, cpu time = %f s\n", fdate(),
             Smoothing.Radius[ismooth], sqrt(Smoothing.Variance[ismooth]), sqrt(Smoothing.TrueVariance[ismooth]),
             tm.per_radius[ismooth]);
  cputime.deriv += tm.deriv;
  cputime.coll += tm.coll;

  if (compute_displacements(1, 0, ScaleDep.z[0]))
    return 1;
  if (Fmax_PDF())        /* histogram on the device, before the context can be released */
    return 1;
  if (finalize_fft())
    return 1;

  cputime.fmax = MPI_Wtime() - cputime.fmax;
  if (!ThisTask)
    printf("[%s] Finishing fmax, total fmax cpu time = %14.6f\n", fdate(), cputime.fmax);
  return 0;
}

/* ---- replaces src/fmax-pfft.c:191-228, 459-560: the FFT work vectors of the host code ---------
 * Used outside the replaced files only by the density special mode (src/pinocchio.c:146-150) and by
 * -DWHITENOISE (src/ReadWhiteNoise.c:161,222).  rvector_fft / cvector_fft stay the host arrays that
 * allocate_fft_vectors hands out (src/allocations.c:380-394); the transform itself runs on the device.
 * Layouts: real [GSlocal_x][N][N], half-complex [N][GSlocal_k_y][N/2+1] (set_one_grid above). */
void write_in_cvector(int ThisGrid, double *restrict vector)
{
  /* GenIC_large leaves delta_k on the device; the reference's leaves it in the host array kdensity[]
   * (src/GenIC.c:384).  The only reader of that array outside the replaced files is special mode 2
   * (src/pinocchio.c:136-168: write_in_cvector(kdensity) -> reverse_transform -> write_density), so it
   * is fetched here, when somebody asks for it, instead of after every GenIC. */
  if (vector == (double *)kdensity[ThisGrid] && pinb200_download_kdensity(pinb, vector))
    pinb_fail("write_in_cvector (kdensity download)");
  memcpy(cvector_fft[ThisGrid], vector, (size_t)MyGrids[ThisGrid].total_local_size_fft * sizeof(double));
}

void write_from_cvector(int ThisGrid, double *restrict vector)
{
  memcpy(vector, cvector_fft[ThisGrid], (size_t)MyGrids[ThisGrid].total_local_size_fft * sizeof(double));
}

void write_in_rvector(int ThisGrid, double *restrict vector)
{
  memcpy(rvector_fft[ThisGrid], vector, (size_t)MyGrids[ThisGrid].total_local_size * sizeof(double));
}

void write_from_rvector(int ThisGrid, double *restrict vector)
{
  memcpy(vector, rvector_fft[ThisGrid], (size_t)MyGrids[ThisGrid].total_local_size * sizeof(double));
}

double forward_transform(int ThisGrid)
{
  double t0 = MPI_Wtime();
  if (pinb200_fft_r2c(pinb, rvector_fft[ThisGrid], (double *)cvector_fft[ThisGrid]))
    pinb_fail("forward_transform");
  return MPI_Wtime() - t0;
}

double reverse_transform(int ThisGrid)
{
  /* includes the 1/N^3 of src/fmax-pfft.c:220-225 */
  double t0 = MPI_Wtime();
  if (pinb200_fft_c2r(pinb, (double *)cvector_fft[ThisGrid], rvector_fft[ThisGrid]))
    pinb_fail("reverse_transform");
  return MPI_Wtime() - t0;
}

/* ---- replaces src/fmax.c:372-506: the DumpProducts file boundary ------------------------------
 * DumpDir/summary (four "%d   # ..." lines: NTasks, RandomSeed, GridSize, sizeof(product_data)),
 * DumpDir/TrueVariance (Nsmooth raw doubles), DumpDir/Task.<rank> (raw products[] of the rank). */
static FILE *dump_open(const char *name, int task, const char *mode)
{
  char fname[LBLENGTH];
  FILE *f;
  if (task >= 0)
    sprintf(fname, "%s%s.%d", params.DumpDir, name, task);
  else
    sprintf(fname, "%s%s", params.DumpDir, name);
  f = fopen(fname, mode);
  if (!f)
    printf("ERROR on Task %d: could not open file %s\n", ThisTask, fname);
  return f;
}

int dump_products(void)
{
  FILE *f;
  int err = 0;
  if (!ThisTask)
  {
    struct stat st;
    if (stat(params.DumpDir, &st))
    {
      printf("Creating directory %s\n", params.DumpDir);
      if (mkdir(params.DumpDir, 0755))
      {
        printf("ERROR IN CREATING DIRECTORY %s (task 0)\n", params.DumpDir);
        err = 1;
      }
    }
    if (!err && (f = dump_open("summary", -1, "w")))
    {
      fprintf(f, "%d   # NTasks\n%d   # random seed\n%d   # grid size\n%d   # length of product_data\n", NTasks,
              params.RandomSeed, params.GridSize[0], (int)sizeof(product_data));
      fclose(f);
    }
    else
      err = 1;
    if (!err && (f = dump_open("TrueVariance", -1, "wb")))
    {
      fwrite(Smoothing.TrueVariance, sizeof(double), Smoothing.Nsmooth, f);
      fclose(f);
    }
    else
      err = 1;
  }
  MPI_Barrier(MPI_COMM_WORLD); /* the directory exists before anybody writes into it */
  if (err || !(f = dump_open("Task", ThisTask, "wb")))
    return 1;
#if !defined(SNAPSHOT) && !defined(RECOMPUTE_DISPLACEMENTS)
  if (pinb)
  {
    /* the records come straight from the device SoA through pinned staging (pinb200_write_products): the file is
       the same, and it does not depend on what the hand-off left in the host products[] */
    pinb200_product_layout L = product_layout();
    fflush(f);
    err = pinb200_write_products(pinb, fileno(f), &L, 0, MyGrids[0].total_local_size);
    fclose(f);
    if (err)
      return pinb_fail("dump_products");
    pinb200_destroy(pinb);
    pinb = NULL;
    return 0;
  }
#endif
  fwrite(products, sizeof(product_data), MyGrids[0].total_local_size, f);
  fclose(f);
  return 0;
}

int read_dumps(void)
{
  FILE *f;
  if (!ThisTask)
  {
    int v[4], want[4], i, bad = 0;
    static const char *what[4] = {"number of tasks", "random seed", "grid size", "length of product_data"};
    char buf[SBLENGTH];
    want[0] = NTasks; want[1] = params.RandomSeed; want[2] = params.GridSize[0]; want[3] = (int)sizeof(product_data);
    if (!(f = dump_open("summary", -1, "r")))
      return 1;
    for (i = 0; i < 4; i++)
      if (!fgets(buf, SBLENGTH, f) || sscanf(buf, "%d", &v[i]) != 1)
        v[i] = -1;
    fclose(f);
    for (i = 0; i < 4; i++)
      if (v[i] != want[i])
      {
        printf("ERROR: the %s in %ssummary does not match - %d vs %d\n", what[i], params.DumpDir, v[i], want[i]);
        bad++;
      }
    if (bad)
      return 1;
    if (!(f = dump_open("TrueVariance", -1, "rb")))
      return 1;
    if (fread(Smoothing.TrueVariance, sizeof(double), Smoothing.Nsmooth, f) != (size_t)Smoothing.Nsmooth)
      printf("WARNING: short read of %sTrueVariance\n", params.DumpDir);
    fclose(f);
  }
  MPI_Bcast(Smoothing.TrueVariance, Smoothing.Nsmooth, MPI_DOUBLE, 0, MPI_COMM_WORLD);
  if (!(f = dump_open("Task", ThisTask, "rb")))
    return 1;
  if (fread(products, sizeof(product_data), MyGrids[0].total_local_size, f) != MyGrids[0].total_local_size)
  {
    printf("ERROR on Task %d: short read of the products dump\n", ThisTask);
    fclose(f);
    return 1;
  }
  fclose(f);
  return 0;
}

/* ---- replaces src/collapse_times.c:780-1346 (-DTABULATED_CT) -----------------------------------------
 * The tables of all smoothing radii are filled in one go on the device (ell_classic or the ELL_SNG
 * ellipsoid integration per table point), or read from params.CTtableFile; the file the reference
 * writes (header + per radius {int ismooth, CT_table}) is written from the downloaded tables. */
#if defined(ELL_SNG) && !defined(TABULATED_CT)
#error "ELL_SNG without TABULATED_CT (one ODE integration per cell and radius) is not provided by the GPU path"
#endif
#if defined(MOD_GRAV_FR) && !defined(ELL_SNG)
#error "MOD_GRAV_FR needs ELL_SNG (the force modification lives in sng_system, src/collapse_times.c:270-311)"
#endif
#ifdef TABULATED_CT
#define SHIM_CT_NBINS_XY 50 /* CT_NBINS_XY, CT_NBINS_D, CT_RANGE_X of src/collapse_times.c:781-787 */
#define SHIM_CT_NBINS_D 100
#define SHIM_CT_RANGE_X 3.5
static int shim_ct_ready = 0;

static int shim_ct_model(void)
{
#ifdef MOD_GRAV_FR
  return PINB200_CT_SNG_FR;
#elif defined(ELL_SNG)
  return PINB200_CT_SNG;
#else
  return PINB200_CT_CLASSIC;
#endif
}

static void shim_ct_write_header(FILE *f, int npoints)
{
  /* write_CTtable_header, src/collapse_times.c:1307-1346 */
  int dummy = shim_ct_model();
  fwrite(&dummy, sizeof(int), 1, f);
  fwrite(&params.Omega0, sizeof(double), 1, f);
  fwrite(&params.OmegaLambda, sizeof(double), 1, f);
  fwrite(&params.Hubble100, sizeof(double), 1, f);
  fwrite(&npoints, sizeof(int), 1, f);
  dummy = SHIM_CT_NBINS_D;
  fwrite(&dummy, sizeof(int), 1, f);
  dummy = SHIM_CT_NBINS_XY;
  fwrite(&dummy, sizeof(int), 1, f);
}

static int shim_ct_check_header(FILE *f, int npoints)
{
  /* check_CTtable_header, src/collapse_times.c:1226-1303 */
  int fail = 0, dummy = 0;
  double fdummy = 0.0;
  if (fread(&dummy, sizeof(int), 1, f) != 1 || dummy != shim_ct_model())
  {
    printf("ERROR: CT table not constructed for this collapse model, %d\n", dummy);
    fail = 1;
  }
  if (fread(&fdummy, sizeof(double), 1, f) != 1 || fabs(fdummy - params.Omega0) > 1.e-10)
  {
    printf("ERROR: CT table constructed for the wrong Omega0, %f in place of %f\n", fdummy, params.Omega0);
    fail = 1;
  }
  if (fread(&fdummy, sizeof(double), 1, f) != 1 || fabs(fdummy - params.OmegaLambda) > 1.e-10)
  {
    printf("ERROR: CT table constructed for the wrong OmegaLambda, %f in place of %f\n", fdummy, params.OmegaLambda);
    fail = 1;
  }
  if (fread(&fdummy, sizeof(double), 1, f) != 1 || fabs(fdummy - params.Hubble100) > 1.e-10)
  {
    printf("ERROR: CT table constructed for the wrong Hubble100, %f in place of %f\n", fdummy, params.Hubble100);
    fail = 1;
  }
  if (fread(&dummy, sizeof(int), 1, f) != 1 || dummy != npoints)
  {
    printf("ERROR: CT table has the wrong size, %d in place of %d\n", dummy, npoints);
    fail = 1;
  }
  if (fread(&dummy, sizeof(int), 1, f) != 1 || dummy != SHIM_CT_NBINS_D)
  {
    printf("ERROR: CT table has the wrong density sampling, %d in place of %d\n", dummy, SHIM_CT_NBINS_D);
    fail = 1;
  }
  if (fread(&dummy, sizeof(int), 1, f) != 1 || dummy != SHIM_CT_NBINS_XY)
  {
    printf("ERROR: CT table has the wrong x and y sampling, %d in place of %d\n", dummy, SHIM_CT_NBINS_XY);
    fail = 1;
  }
  return fail;
}

/* all radii at once; onlycompute as in initialize_collapse_times(ismooth, onlycompute) */
static int shim_collapse_tables(int onlycompute)
{
  const int ns = Smoothing.Nsmooth, npoints = SHIM_CT_NBINS_D * SHIM_CT_NBINS_XY * SHIM_CT_NBINS_XY;
  int ismooth, fail = 0;
  pinb200_ct_desc d;
  double *d_in = (double *)malloc(ns * sizeof(double));
  double *fr_size = (double *)malloc(ns * sizeof(double));
  double *tables = NULL;
  const int from_file = strcmp(params.CTtableFile, "none") && !onlycompute;

  memset(&d, 0, sizeof d);
  d.model = shim_ct_model();
  d.nbins_d = SHIM_CT_NBINS_D;
  d.nbins_xy = SHIM_CT_NBINS_XY;
  d.range_x = SHIM_CT_RANGE_X;
  d.delta_vector = NULL; /* the reference's compiled sampling */
#ifdef ELL_SNG
  /* OmegaMatter(z), OmegaLambda(z) of src/cosmo.c:1675-1718 as closed forms of a cosmological constant */
  if (!params.simpleLambda)
  {
    printf("ERROR on task %d: the GPU path integrates ELL_SNG for a cosmological constant only\n", ThisTask);
    return 1;
  }
  d.omega0 = params.Omega0;
  d.omega_lambda = params.OmegaLambda;
#ifdef NORADIATION
  d.omega_rad = 0.0; /* OMEGARAD_H2, src/cosmo.c:36-40 */
#else
  d.omega_rad = 4.2e-5 / params.Hubble100 / params.Hubble100;
#endif
  d.omega_k = 1.0 - params.Omega0 - params.OmegaLambda - d.omega_rad;
  {
    /* the closed form must be the host cosmology's own E(z) (no READ_HUBBLE_TABLE, no tabulated EoS) */
    const double z = 1.0, e2 = d.omega_rad * pow(1. + z, 4.) + d.omega0 * pow(1. + z, 3.) + d.omega_k * pow(1. + z, 2.) + d.omega_lambda;
    const double om = d.omega0 * pow(1. + z, 3.) / e2;
    if (fabs(om - OmegaMatter(z)) > 1.e-12)
    {
      printf("ERROR on task %d: OmegaMatter(z) of the host cosmology is not that of a cosmological constant\n", ThisTask);
      return 1;
    }
  }
#endif
  for (ismooth = 0; ismooth < ns; ismooth++)
  {
    /* D_in of ell_sng, src/collapse_times.c:345-353 */
#ifdef SCALE_DEPENDENT
    d_in[ismooth] = GrowingMode(1. / 1.e-5 - 1., params.k_for_GM / Smoothing.Radius[ismooth] * params.InterPartDist);
#else
    d_in[ismooth] = GrowingMode(1. / 1.e-5 - 1., 1. / Smoothing.Radius[ismooth]);
#endif
    /* ode_param of ell_sng, src/collapse_times.c:362-372 */
    fr_size[ismooth] = (ismooth < ns - 1 || ns == 1) ? Smoothing.Radius[ismooth] : Smoothing.Radius[ismooth - 1];
  }
#ifdef MOD_GRAV_FR
  d.fr0 = FR0;
  d.h_over_c = H_over_c;
  d.fr_size = fr_size;
#endif

  if (from_file)
  {
    /* Task 0 reads, everybody gets the tables (src/collapse_times.c:937-963) */
    tables = (double *)malloc((size_t)ns * npoints * sizeof(double));
    if (!ThisTask)
    {
      FILE *f = fopen(params.CTtableFile, "r");
      if (!f)
      {
        printf("ERROR: cannot open CTtableFile %s\n", params.CTtableFile);
        fail = 1;
      }
      else
      {
        fail = shim_ct_check_header(f, npoints);
        for (ismooth = 0; ismooth < ns && !fail; ismooth++)
        {
          int dummy;
          if (fread(&dummy, sizeof(int), 1, f) != 1 || fread(tables + (size_t)ismooth * npoints, sizeof(double), npoints, f) != (size_t)npoints)
          {
            printf("ERROR: short read of CTtableFile %s at smoothing radius %d\n", params.CTtableFile, ismooth);
            fail = 1;
          }
        }
        fclose(f);
      }
    }
    MPI_Bcast(&fail, sizeof(int), MPI_BYTE, 0, MPI_COMM_WORLD);
    if (fail)
      return 1;
    MPI_Bcast(tables, ns * npoints, MPI_DOUBLE, 0, MPI_COMM_WORLD);
  }

  if (pinb200_set_collapse_tables(pinb, &d, Smoothing.Variance, d_in, tables))
    return pinb_fail("pinb200_set_collapse_tables");
  free(d_in);
  free(fr_size);
  if (tables)
    free(tables);

  if (!from_file && !ThisTask)
  {
    /* the table file of src/collapse_times.c:991-1033 (binary form) */
    char fname[LBLENGTH];
    FILE *f;
    double *t = (double *)malloc((size_t)npoints * sizeof(double));
    if (onlycompute)
      strcpy(fname, params.CTtableFile);
    else
      sprintf(fname, "pinocchio.%s.CTtable.out", params.RunFlag);
    f = fopen(fname, "w");
    if (!f)
    {
      printf("ERROR: cannot write %s\n", fname);
      return 1;
    }
    shim_ct_write_header(f, npoints);
    for (ismooth = 0; ismooth < ns; ismooth++)
    {
      if (pinb200_download_collapse_table(pinb, ismooth, t))
        return pinb_fail("pinb200_download_collapse_table");
      fwrite(&ismooth, sizeof(int), 1, f);
      fwrite(t, sizeof(double), npoints, f);
    }
    fclose(f);
    free(t);
  }
  shim_ct_ready = 1;
  return 0;
}

/* external symbols of src/pinocchio.h:546-549: the special mode `pinocchio.x parameter_file 1`
 * (src/pinocchio.c:97-125) calls this once per radius; all tables are made at the first call */
int initialize_collapse_times(int ismooth, int onlycompute)
{
  if (ismooth == 0 || !shim_ct_ready)
    return shim_collapse_tables(onlycompute);
  return 0;
}

int reset_collapse_times(int ismooth)
{
  (void)ismooth;
  return 0;
}
#endif /* TABULATED_CT */

char *fdate(void)
{
  /* identical output format to src/fmax.c:261-289 */
  time_t current_time = time(NULL);
  char *string = ctime(&current_time);
  int n;
  for (n = 0; n < 10; n++) *(date_string + n) = *(string + n);
  for (n = 10; n < 15; n++) *(date_string + n) = *(string + n + 9);
  for (n = 10; n < 19; n++) *(date_string + n + 5) = *(string + n);
  *(date_string + 24) = '\0';
  return date_string;
}
