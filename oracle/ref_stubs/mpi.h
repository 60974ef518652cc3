/* Declaration-only stand-in for MPI (absent from this image); oracle/ref_harness.c implements
 * the calls for one task. */
#ifndef PINB_REFSTUB_MPI_H
#define PINB_REFSTUB_MPI_H
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef struct { int s; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_DOUBLE 2
#define MPI_FLOAT 3
#define MPI_INT 4
#define MPI_UNSIGNED 5
#define MPI_UNSIGNED_LONG_LONG 6
#define MPI_SUM 1
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm);
int MPI_Comm_free(MPI_Comm*);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
#endif
