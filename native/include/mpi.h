/* shipyard mini-MPI: the subset of MPI the recipe workloads use (mpiBench-style collective
 * sweeps, OSU-style latency loops, HPCG dot products + halo exchange), implemented by
 * libshipyard_mpi over the shipyard collectives (sm_100a kernels for device buffers, host
 * shared memory for host buffers).  Ranks come from the task runner's environment; there is
 * no mpirun.  Handles are small integers private to this implementation. */
#ifndef SHIPYARD_MPI_H
#define SHIPYARD_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int _count; } MPI_Status;

#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-2)
#define MPI_REQUEST_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_MAX_PROCESSOR_NAME 128
#define MPI_MAX_ERROR_STRING 256
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

enum { MPI_DATATYPE_NULL = 0, MPI_CHAR, MPI_SIGNED_CHAR, MPI_UNSIGNED_CHAR, MPI_BYTE, MPI_SHORT, MPI_UNSIGNED_SHORT,
       MPI_INT, MPI_UNSIGNED, MPI_LONG, MPI_UNSIGNED_LONG, MPI_LONG_LONG, MPI_LONG_LONG_INT, MPI_UNSIGNED_LONG_LONG,
       MPI_FLOAT, MPI_DOUBLE, MPI_INT32_T, MPI_INT64_T, MPI_UINT8_T, MPI_UINT32_T, MPI_UINT64_T,
       MPIX_BFLOAT16, MPIX_FLOAT16 };
enum { MPI_OP_NULL = 0, MPI_SUM, MPI_MAX, MPI_MIN, MPI_PROD };

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Get_processor_name(char* name, int* resultlen);
int MPI_Error_string(int errorcode, char* string, int* resultlen);
int MPI_Type_size(MPI_Datatype dt, int* size);
int MPI_Get_count(const MPI_Status* status, MPI_Datatype dt, int* count);
double MPI_Wtime(void);
double MPI_Wtick(void);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Reduce_scatter_block(const void* sendbuf, void* recvbuf, int recvcount, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, int recvcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, const int* recvcounts, const int* displs, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, int recvcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sdt, void* recvbuf, const int* recvcounts, const int* rdispls, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, int recvcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, const int* recvcounts, const int* displs, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, int recvcount, MPI_Datatype rdt, int root, MPI_Comm comm);

int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sdt, int dest, int sendtag, void* recvbuf, int recvcount,
                 MPI_Datatype rdt, int source, int recvtag, MPI_Comm comm, MPI_Status* status);

/* shipyard extensions */
int MPIX_Query_shipyard_transport(char* name, int len);   /* "stub" | "p2p" | "nvls" for device buffers */
void* MPIX_Sym_alloc(size_t bytes);
void* MPIX_Device_stream(void);                           /* cudaStream_t the device-buffer collectives run on */
int MPIX_Set_device_async(int on);                        /* 1: device-buffer collectives return after the enqueue */
int MPIX_Device_sync(void);                               /* synchronise that stream and check the collective watchdog */
void* MPIX_Shipyard_comm(void);                           /* the sy_comm* behind the device-buffer collectives */                        /* symmetric device allocation (zero-copy collectives) */

#ifdef __cplusplus
}
#endif
#endif
