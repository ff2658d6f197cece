/* TEST / BENCH INFRASTRUCTURE ONLY.  See mpi.h. */
#define _GNU_SOURCE
#include "mpi.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define MAXP 64

typedef struct {
    pthread_barrier_t bar;
    int nprocs;
    size_t arena_bytes;
    long long counts[MAXP * MAXP]; /* per-collective metadata published by every rank */
    long long displs[MAXP * MAXP];
    size_t offset[MAXP];           /* where each rank's payload starts in the arena */
    size_t length[MAXP];
} shared_t;

static shared_t* g_sh = NULL;
static unsigned char* g_arena = NULL;
static int g_rank = 0;

static size_t tsize(MPI_Datatype t) { return t == MPI_INT ? sizeof(int) : sizeof(float); }
static void sync_all(void) { pthread_barrier_wait(&g_sh->bar); }

int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = g_sh ? g_sh->nprocs : 1; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { (void)c; sync_all(); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; _exit(code ? code : 1); }

static void need(size_t bytes) {
    if (bytes > g_sh->arena_bytes) {
        fprintf(stderr, "mpi_standin: collective needs %zu bytes, arena has %zu\n", bytes, g_sh->arena_bytes);
        _exit(3);
    }
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
    (void)comm;
    const size_t n = (size_t)count * tsize(type);
    need(n);
    if (g_rank == root) memcpy(g_arena, buf, n);
    sync_all();
    if (g_rank != root) memcpy(buf, g_arena, n);
    sync_all();
    return MPI_SUCCESS;
}

int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype st, void* recvbuf, int recvcount,
                 MPI_Datatype rt, int root, MPI_Comm comm) {
    (void)comm; (void)rt;
    const int P = g_sh->nprocs;
    const size_t es = tsize(st);
    if (g_rank == root) {
        size_t off = 0;
        for (int r = 0; r < P; ++r) {
            g_sh->offset[r] = off;
            g_sh->length[r] = (size_t)sendcounts[r] * es;
            need(off + g_sh->length[r]);
            memcpy(g_arena + off, (const unsigned char*)sendbuf + (size_t)displs[r] * es, g_sh->length[r]);
            off += g_sh->length[r];
        }
    }
    sync_all();
    {
        size_t n = (size_t)recvcount * es;
        if (n > g_sh->length[g_rank]) n = g_sh->length[g_rank];
        memcpy(recvbuf, g_arena + g_sh->offset[g_rank], n);
    }
    sync_all();
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype st, void* recvbuf, const int* recvcounts, const int* displs,
                MPI_Datatype rt, int root, MPI_Comm comm) {
    (void)comm; (void)rt;
    const int P = g_sh->nprocs;
    const size_t es = tsize(st);
    g_sh->length[g_rank] = (size_t)sendcount * es;
    sync_all();
    {
        size_t off = 0;
        for (int r = 0; r < g_rank; ++r) off += g_sh->length[r];
        need(off + g_sh->length[g_rank]);
        memcpy(g_arena + off, sendbuf, g_sh->length[g_rank]);
    }
    sync_all();
    if (g_rank == root) {
        size_t off = 0;
        for (int r = 0; r < P; ++r) {
            size_t n = (size_t)recvcounts[r] * es;
            if (n > g_sh->length[r]) n = g_sh->length[r];
            memcpy((unsigned char*)recvbuf + (size_t)displs[r] * es, g_arena + off, n);
            off += g_sh->length[r];
        }
    }
    sync_all();
    return MPI_SUCCESS;
}

int MPI_Alltoallv(const void* sendbuf, const int* sendcounts, const int* sdispls, MPI_Datatype st, void* recvbuf,
                  const int* recvcounts, const int* rdispls, MPI_Datatype rt, MPI_Comm comm) {
    (void)comm; (void)rt;
    const int P = g_sh->nprocs;
    const size_t es = tsize(st);
    size_t mine = 0;
    for (int d = 0; d < P; ++d) {
        g_sh->counts[g_rank * MAXP + d] = sendcounts[d];
        mine += (size_t)sendcounts[d] * es;
    }
    g_sh->length[g_rank] = mine;
    sync_all();
    {   /* rank r's outgoing blocks are packed contiguously, in destination order, after ranks 0..r-1 */
        size_t off = 0;
        for (int r = 0; r < g_rank; ++r) off += g_sh->length[r];
        need(off + mine);
        for (int d = 0; d < P; ++d) {
            const size_t n = (size_t)sendcounts[d] * es;
            memcpy(g_arena + off, (const unsigned char*)sendbuf + (size_t)sdispls[d] * es, n);
            off += n;
        }
    }
    sync_all();
    {
        size_t base = 0;
        for (int s = 0; s < P; ++s) {
            size_t off = base;
            for (int d = 0; d < g_rank; ++d) off += (size_t)g_sh->counts[s * MAXP + d] * es;
            size_t n = (size_t)g_sh->counts[s * MAXP + g_rank] * es;
            const size_t cap = (size_t)recvcounts[s] * es;
            if (n > cap) n = cap;
            memcpy((unsigned char*)recvbuf + (size_t)rdispls[s] * es, g_arena + off, n);
            base += g_sh->length[s];
        }
    }
    sync_all();
    return MPI_SUCCESS;
}

int mpi_standin_launch(int nprocs, size_t arena_bytes, void (*fn)(void*), void* arg) {
    if (nprocs < 1 || nprocs > MAXP) return -1;
    g_sh = (shared_t*)mmap(NULL, sizeof(shared_t), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    g_arena = (unsigned char*)mmap(NULL, arena_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (g_sh == MAP_FAILED || g_arena == MAP_FAILED) return -2;
    g_sh->nprocs = nprocs;
    g_sh->arena_bytes = arena_bytes;
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&g_sh->bar, &at, (unsigned)nprocs);
    pid_t kids[MAXP];
    for (int r = 1; r < nprocs; ++r) {
        pid_t pid = fork();
        if (pid < 0) return -3;
        if (pid == 0) {
            g_rank = r;
            fn(arg);
            fflush(NULL);
            _exit(0);
        }
        kids[r] = pid;
    }
    g_rank = 0;
    fn(arg);
    int bad = 0;
    for (int r = 1; r < nprocs; ++r) {
        int st = 0;
        waitpid(kids[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
    }
    munmap(g_arena, arena_bytes);
    munmap(g_sh, sizeof(shared_t));
    g_sh = NULL;
    return bad ? -4 : 0;
}
