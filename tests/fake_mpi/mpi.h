// Just enough of <mpi.h> to compile and RUN csrc/glb/mpi/context.cc without an MPI
// installation: the "world" is a set of threads in one process, each of which announces its
// rank with fake_mpi_set_rank() before touching MPI. Implemented in fake_mpi.cc.
#pragma once

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_INT 1
#define MPI_BYTE 2
#define MPI_MAX 1
#define MPI_THREAD_MULTIPLE 3

#ifdef __cplusplus
extern "C" {
#endif
void fake_mpi_world(int size);   // once, before the rank threads start
void fake_mpi_set_rank(int rank); // per rank thread
int fake_mpi_live_comms(void);    // dup'ed communicators not yet freed

int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* out);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* send, int scount, MPI_Datatype sdt, void* recv, int rcount, MPI_Datatype rdt,
                  MPI_Comm comm);
#ifdef __cplusplus
}
#endif
