/* TEST INFRASTRUCTURE (oracle/_ref/libshim_host.so): lets the HOST-side functions of the drop-in
 * shim (shim/fmax_b200.c: dump_products, read_dumps, set_one_grid, the work-vector copies) run in
 * this container, linked exactly as a PINOCCHIO build would link them -- against the reference's
 * own globals (src/variables.c, compiled from where it lies) and libpinb200.so -- so that they
 * can be compared with the reference's functions of the same name living in libpinocchio_ref.so.
 * One-task MPI and the few cosmology symbols the shim references are stand-ins; no GPU entry
 * point is called from here.  Nothing in the product links or loads this file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pinocchio.h"

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static size_t mpi_size(MPI_Datatype t) { return (t == MPI_DOUBLE || t == MPI_UNSIGNED_LONG_LONG) ? 8 : (t == MPI_BYTE ? 1 : 4); }
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int r, MPI_Comm c) { (void)b; (void)n; (void)t; (void)r; (void)c; return MPI_SUCCESS; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) {
  (void)o; (void)root; (void)c;
  memcpy(r, s, n * mpi_size(t));
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { return MPI_Reduce(s, r, n, t, o, 0, c); }
int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c;
  memcpy(r, s, n * mpi_size(t));
  return MPI_SUCCESS;
}
/* cosmology symbols the shim references (src/cosmo.c needs GSL): never called by these tests */
double PowerSpectrum(double k) { (void)k; abort(); }
double GrowingMode(double z, double k) { (void)z; (void)k; abort(); }
double GrowingMode_2LPT(double z, double k) { (void)z; (void)k; abort(); }
double GrowingMode_3LPT_1(double z, double k) { (void)z; (void)k; abort(); }
double GrowingMode_3LPT_2(double z, double k) { (void)z; (void)k; abort(); }

int dump_products(void);
int read_dumps(void);
int set_one_grid(int);

/* the reference globals these functions read: one grid of N^3 on `ntasks` tasks, this one = `task` */
int host_setup(int N, int task, int ntasks, int seed, int nsmooth, const char* dumpdir) {
  ThisTask = task;
  NTasks = ntasks;
  Ngrids = 1;
  memset(&params, 0, sizeof(params));
  for (int i = 0; i < 3; i++) params.GridSize[i] = N;
  params.RandomSeed = seed;
  snprintf(params.DumpDir, SBLENGTH, "%s", dumpdir);
  free(MyGrids);
  MyGrids = calloc(1, sizeof(grid_data));
  for (int i = 0; i < 3; i++) MyGrids[0].GSglobal[i] = N;
  MyGrids[0].Ntotal = (unsigned long long)N * N * N;
  MyGrids[0].BoxSize = (double)N;
  if (set_one_grid(0)) return 1;
  Smoothing.Nsmooth = nsmooth;
  free(Smoothing.TrueVariance);
  Smoothing.TrueVariance = calloc(nsmooth, sizeof(double));
  free(products);
  products = calloc(MyGrids[0].total_local_size, sizeof(product_data));
  return 0;
}
long host_local_cells(void) { return (long)MyGrids[0].total_local_size; }
int host_sizeof_product(void) { return (int)sizeof(product_data); }
/* slab geometry as set_one_grid left it: GSlocal[3], GSstart[3], GSlocal_k[3], GSstart_k[3], total_local_size_fft */
int host_geometry(long* out) {
  for (int i = 0; i < 3; i++) {
    out[i] = MyGrids[0].GSlocal[i];
    out[3 + i] = MyGrids[0].GSstart[i];
    out[6 + i] = MyGrids[0].GSlocal_k[i];
    out[9 + i] = MyGrids[0].GSstart_k[i];
  }
  out[12] = MyGrids[0].total_local_size_fft;
  return 0;
}
int host_set_products(const void* rec, const double* tv) {
  memcpy(products, rec, (size_t)MyGrids[0].total_local_size * sizeof(product_data));
  memcpy(Smoothing.TrueVariance, tv, Smoothing.Nsmooth * sizeof(double));
  return 0;
}
int host_get_products(void* rec, double* tv) {
  memcpy(rec, products, (size_t)MyGrids[0].total_local_size * sizeof(product_data));
  memcpy(tv, Smoothing.TrueVariance, Smoothing.Nsmooth * sizeof(double));
  return 0;
}
int host_dump_products(void) { return dump_products(); }
int host_read_dumps(void) { return read_dumps(); }
