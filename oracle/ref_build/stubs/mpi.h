/* Declaration-only stand-in for <mpi.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's hot-path operator translation units
 * include <mpi.h> transitively (Tools.h) although the operator arithmetic never
 * calls MPI.  This header only lets those translation units *compile* from where
 * they lie under /root/reference; every function is an inert inline no-op that
 * reports a single rank.  Nothing from the reference is copied here.
 */
#ifndef SB200_STUB_MPI_H
#define SB200_STUB_MPI_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Group;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_PROC_NULL (-2)
#define MPI_SUCCESS 0
#define MPI_TAG_UB 1
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_THREAD_MULTIPLE 3
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_FUNNELED 1
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_MINLOC 5
#define MPI_MAXLOC 6
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_UNSIGNED 3
#define MPI_UINT 3
#define MPI_LONG 4
#define MPI_UNSIGNED_LONG 5
#define MPI_DOUBLE 6
#define MPI_FLOAT 7
#define MPI_SHORT 8
#define MPI_UNSIGNED_LONG_LONG 9
#define MPI_LONG_LONG_INT 10
#define MPI_DOUBLE_INT 11
#define MPI_BYTE 12
#define MPI_UNSIGNED_SHORT 13
#define MPI_UINT64_T 14
#define MPI_2INT 15
#define MPI_C_BOOL 16
#define MPI_CXX_BOOL 17
#define MPI_LONG_DOUBLE 18
#define MPI_ORDER_C 0
#define MPI_MAX_PROCESSOR_NAME 64
#ifdef __cplusplus
template<class... A> static inline int MPI_Comm_rank( MPI_Comm, int *r, A... ) { *r = 0; return 0; }
template<class... A> static inline int MPI_Comm_size( MPI_Comm, int *s, A... ) { *s = 1; return 0; }
static inline double MPI_Wtime() { return 0.; }
#define SB200_MPI_NOOP(name) template<class... A> static inline int name( A... ) { return 0; }
SB200_MPI_NOOP(MPI_Barrier) SB200_MPI_NOOP(MPI_Comm_free) SB200_MPI_NOOP(MPI_Comm_get_attr)
SB200_MPI_NOOP(MPI_Isend) SB200_MPI_NOOP(MPI_Irecv) SB200_MPI_NOOP(MPI_Send) SB200_MPI_NOOP(MPI_Recv)
SB200_MPI_NOOP(MPI_Wait) SB200_MPI_NOOP(MPI_Waitall) SB200_MPI_NOOP(MPI_Test)
SB200_MPI_NOOP(MPI_Reduce) SB200_MPI_NOOP(MPI_Allreduce) SB200_MPI_NOOP(MPI_Bcast) SB200_MPI_NOOP(MPI_Gather)
SB200_MPI_NOOP(MPI_Allgather) SB200_MPI_NOOP(MPI_Allgatherv) SB200_MPI_NOOP(MPI_Gatherv) SB200_MPI_NOOP(MPI_Scatter)
SB200_MPI_NOOP(MPI_Scatterv) SB200_MPI_NOOP(MPI_Sendrecv) SB200_MPI_NOOP(MPI_Ssend) SB200_MPI_NOOP(MPI_Issend)
SB200_MPI_NOOP(MPI_Type_commit) SB200_MPI_NOOP(MPI_Type_free) SB200_MPI_NOOP(MPI_Type_contiguous)
SB200_MPI_NOOP(MPI_Type_vector) SB200_MPI_NOOP(MPI_Type_create_struct) SB200_MPI_NOOP(MPI_Type_create_subarray)
SB200_MPI_NOOP(MPI_Type_create_hindexed) SB200_MPI_NOOP(MPI_Type_create_hvector) SB200_MPI_NOOP(MPI_Get_address)
SB200_MPI_NOOP(MPI_Init) SB200_MPI_NOOP(MPI_Init_thread) SB200_MPI_NOOP(MPI_Finalize) SB200_MPI_NOOP(MPI_Abort)
SB200_MPI_NOOP(MPI_Comm_split) SB200_MPI_NOOP(MPI_Comm_dup) SB200_MPI_NOOP(MPI_Comm_group) SB200_MPI_NOOP(MPI_Group_incl)
SB200_MPI_NOOP(MPI_Comm_create) SB200_MPI_NOOP(MPI_Get_count) SB200_MPI_NOOP(MPI_Get_processor_name)
SB200_MPI_NOOP(MPI_Exscan) SB200_MPI_NOOP(MPI_Scan) SB200_MPI_NOOP(MPI_Iprobe) SB200_MPI_NOOP(MPI_Probe)
SB200_MPI_NOOP(MPI_Testall) SB200_MPI_NOOP(MPI_Waitany) SB200_MPI_NOOP(MPI_Cancel) SB200_MPI_NOOP(MPI_Request_free)
#endif
#endif
