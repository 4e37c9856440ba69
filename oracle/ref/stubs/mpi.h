/* Single-rank stand-in for <mpi.h>, only so that the UNMODIFIED reference sources under
 * /root/reference compile into the CPU oracle (oracle/_ref).  Test infrastructure only.
 * A datatype is encoded as its size in bytes; every collective degenerates to a copy. */
#ifndef G2_STUB_MPI_H
#define G2_STUB_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_BYTE   1
#define MPI_CHAR   1
#define MPI_INT    4
#define MPI_FLOAT  4
#define MPI_DOUBLE 8
#define MPI_LONG   8
#define MPI_SUM 0
#define MPI_MIN 1
#define MPI_MAX 2
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int err);
int MPI_Comm_rank(MPI_Comm c, int *r);
int MPI_Comm_size(MPI_Comm c, int *s);
int MPI_Barrier(MPI_Comm c);
double MPI_Wtime(void);
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c);
int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displ, MPI_Datatype tr, MPI_Comm c);
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr, int src, int rtag, MPI_Comm c, MPI_Status *st);
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Send(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
#endif
