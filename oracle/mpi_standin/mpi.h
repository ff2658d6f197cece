/* TEST / BENCH INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Single-node stand-in for the handful of MPI calls the reference's MPI backend makes
 * (/root/reference/fft/fft_mpi.cpp: Comm_rank, Comm_size, Bcast, Scatterv, Alltoallv, Gatherv), so that
 * fft_mpi.cpp compiles UNMODIFIED and runs as P forked processes on one host (the image has no MPI
 * runtime: no mpic++, mpirun or mpi.h).  Ranks are processes created by mpi_standin_launch();
 * collectives copy through one anonymous shared mapping and a process-shared barrier.
 * SURVEY.md 8(f) rank 4. */
#ifndef FDR_MPI_STANDIN_H
#define FDR_MPI_STANDIN_H
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_FLOAT 104
#define MPI_SUCCESS 0

#ifdef __cplusplus
extern "C" {
#endif
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype sendtype, void* recvbuf,
                  const int* recvcounts, const int* rdispls, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int code);

/* Forks nprocs-1 children; every process (parent = rank 0) then runs fn(arg).  Children _exit when fn
 * returns; the parent waits for them and returns 0 if all exited cleanly.  arena_bytes must cover the
 * largest payload any single collective moves in total. */
int mpi_standin_launch(int nprocs, size_t arena_bytes, void (*fn)(void*), void* arg);
#ifdef __cplusplus
}
#endif
#endif
