/* One-task MPI for oracle/_ref/pinocchio_ref.x (MPI is absent from this image): declarations of
 * the calls the reference makes; oracle/ref_full/mini_mpi.c implements them for a single task.
 * TEST INFRASTRUCTURE. */
#ifndef PINB_REFFULL_MPI_H
#define PINB_REFFULL_MPI_H
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_DOUBLE 2
#define MPI_FLOAT 3
#define MPI_INT 4
#define MPI_UNSIGNED 5
#define MPI_UNSIGNED_LONG_LONG 6
#define MPI_CHAR 7
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 9
#define MPI_LONG_LONG 10
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_THREAD_FUNNELED 1
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)-1)
int MPI_Init_thread(int*, char***, int, int*);
int MPI_Init(int*, char***);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_free(MPI_Comm*);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
#endif
