/* Minimal MPI subset for building the PanguLU reference host code in containers
 * that ship no MPI (test infrastructure only -- never part of the product path).
 *
 * Only the calls the reference issues are declared (grep MPI_ over
 * /root/reference/src and examples/example.c): Init_thread, Finalize, Comm_rank,
 * Comm_size, Barrier, Bcast, Send, Recv, Isend, Irecv, Wait, Waitall, Iprobe,
 * Probe, Get_count, Reduce, Allreduce.  The implementation lives in
 * mpi_local.c: rank count 1 by default; with PANGULU_SHIM_NP>1 in the
 * environment the ranks are forked processes that talk through a POSIX
 * shared-memory mailbox (see mpi_local.c).
 */
#ifndef PANGULU_ORACLE_MPI_SHIM_H
#define PANGULU_ORACLE_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value = element size in bytes */
typedef int MPI_Op;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; long long _count; } MPI_Status;
typedef struct { int kind; void *buf; long long bytes; int peer; int tag; int done; } *MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_THREAD_MULTIPLE 3
#define MPI_SUM 1

#define MPI_CHAR 1
#define MPI_SHORT 2
#define MPI_UNSIGNED_SHORT 2
#define MPI_UINT16_T 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_UINT32_T 4
#define MPI_FLOAT 4
#define MPI_LONG_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_UINT64_T 8
#define MPI_DOUBLE 8
#define MPI_C_FLOAT_COMPLEX 8
#define MPI_DOUBLE_COMPLEX 16

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Iprobe(int src, int tag, MPI_Comm comm, int *flag, MPI_Status *st);
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count);
int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
