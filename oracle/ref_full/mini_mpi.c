/* TEST INFRASTRUCTURE: MPI for exactly one task, so that the whole reference program
 * (oracle/_ref/pinocchio_ref.x) runs in an image without MPI.  Collectives copy the send buffer,
 * point-to-point calls cannot occur with one task and abort. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mpi.h>
#include <pfft.h>
#include <fftw3-mpi.h>

static size_t tsize(MPI_Datatype t) {
  switch (t) {
    case MPI_DOUBLE: case MPI_UNSIGNED_LONG_LONG: case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_LONG_LONG: return 8;
    case MPI_FLOAT: case MPI_INT: case MPI_UNSIGNED: return 4;
    default: return 1;
  }
}
static void one_task_only(const char* what) {
  fprintf(stderr, "oracle/ref_full: %s called in a one-task run\n", what);
  abort();
}
int MPI_Init_thread(int* c, char*** v, int req, int* prov) { (void)c; (void)v; if (prov) *prov = req; return MPI_SUCCESS; }
int MPI_Init(int* c, char*** v) { (void)c; (void)v; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; fflush(stdout); exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* c) { (void)c; return MPI_SUCCESS; }
double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)op; (void)root; (void)c;
  if (s != MPI_IN_PLACE && s != r) memmove(r, s, (size_t)n * tsize(t));
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return MPI_Reduce(s, r, n, t, op, 0, c); }
int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c;
  if (s != MPI_IN_PLACE && s != r) memmove(r, s, (size_t)n * tsize(t));
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void* s, int n, MPI_Datatype t, void* r, const int* rc, const int* displs, MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)rt; (void)c;
  if (s != MPI_IN_PLACE) memmove((char*)r + (displs ? (size_t)displs[0] * tsize(t) : 0), s, (size_t)n * tsize(t));
  return MPI_SUCCESS;
}
int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c;
  one_task_only("MPI_Send");
  return 1;
}
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)st;
  one_task_only("MPI_Recv");
  return 1;
}

/* process mesh and thread set-up of PFFT / FFTW (src/initialization.c:176-203) */
int pfft_create_procmesh(int rnk, MPI_Comm comm, const int* np, MPI_Comm* comm_cart) {
  (void)rnk; (void)np;
  *comm_cart = comm;
  return 0;
}
void pfft_plan_with_nthreads(int n) { (void)n; }
int fftw_init_threads(void) { return 1; }
void fftw_mpi_init(void) {}
void fftw_plan_with_nthreads(int n) { (void)n; }
void fftw_cleanup_threads(void) {}
void fftw_mpi_cleanup(void) {}
