/* The MPI host program of INTEGRATION.md section 3, complete: what a maintainer of the reference's mpi.cpp:95-111 would write
 * to restore ONE image row-sharded over `world` ranks through the C ABI (include/fdr_b200.h) -- one fdr_shard per rank, the
 * slab handles exchanged once over MPI, then one fdr_shard_restore_rows per image; no MPI call on the per-image path.
 *
 * TEST PROGRAM: it is compiled against the single-node MPI stand-in (oracle/mpi_standin: ranks are forked processes), and every
 * rank uses CUDA device (rank % device_count), so that a one-GPU box can run it; with a real MPI and one GPU per rank only
 * main() changes (MPI_Init instead of mpi_standin_launch, MPI_Allgather instead of the Bcast loop).
 *
 *   shard_mpi_host <ranks> <rows> <cols> <psf_len> <psf_angle> <images> <out.bin>
 * Writes the restored u8 BGR image [rows][cols][3] of the LAST synthetic image (seed 0xF17E0004) to out.bin. */
#define _GNU_SOURCE
#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fdr_b200.h"
#include "mpi.h"

#define MAXRANKS 8

struct cfg {
    int rows, cols, psf_len, images;
    double psf_angle;
    const char* out_path;
};

#define CHECK_FDR(call)                                                                              \
    do {                                                                                             \
        if ((call) != 0) {                                                                           \
            fprintf(stderr, "Error: %s:%d, %s\n", __FILE__, __LINE__, fdr_last_error());             \
            MPI_Abort(MPI_COMM_WORLD, 1);                                                            \
            _exit(1);                                                                                \
        }                                                                                            \
    } while (0)
#define CHECK_CUDA(call)                                                                             \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "Error: %s:%d, %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));       \
            MPI_Abort(MPI_COMM_WORLD, 1);                                                            \
            _exit(1);                                                                                \
        }                                                                                            \
    } while (0)

static void rank_main(void* arg) {
    const struct cfg* c = (const struct cfg*)arg;
    int rank = 0, world = 1, ndev = 0;
    alarm(100); /* test program: a rank that is stuck (a peer died) must not keep a GPU busy */
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &world);
    CHECK_FDR(fdr_device_count(&ndev));
    if (ndev < 1) {
        fprintf(stderr, "Error: no CUDA device (there is no CPU fallback)\n");
        _exit(3);
    }
    const int device = rank % ndev;
    CHECK_CUDA(cudaSetDevice(device));

    fdr_shard* sh = NULL;
    CHECK_FDR(fdr_shard_create(&sh, c->rows, c->cols, 3, rank, world, device));
    int first_row = 0, n_rows = 0;
    CHECK_FDR(fdr_shard_geometry(sh, &first_row, &n_rows, NULL, NULL, NULL));

    /* slab handles: 64 bytes per rank (MPI_Allgather with a real MPI; the stand-in has Bcast) */
    void* slab = NULL;
    size_t slab_bytes = 0;
    CHECK_FDR(fdr_shard_local_slab(sh, &slab, &slab_bytes));
    unsigned char all[MAXRANKS][64];
    CHECK_FDR(fdr_ipc_export(slab, all[rank]));
    for (int r = 0; r < world; ++r) MPI_Bcast(all[r], 16, MPI_INT, r, MPI_COMM_WORLD);
    void* peers[MAXRANKS];
    for (int r = 0; r < world; ++r) {
        if (r == rank)
            peers[r] = slab;
        else
            CHECK_FDR(fdr_ipc_open(all[r], &peers[r]));
    }
    CHECK_FDR(fdr_shard_set_peers(sh, peers));
    CHECK_FDR(fdr_shard_set_psf_motion(sh, c->psf_len, c->psf_angle, 0.01f));
    MPI_Barrier(MPI_COMM_WORLD); /* fence: the Wiener build used the column slab as scratch */

    const size_t row_bytes = (size_t)c->cols * 3;
    unsigned char *d_in = NULL, *d_out = NULL;
    CHECK_CUDA(cudaMalloc((void**)&d_in, row_bytes * (n_rows > 0 ? n_rows : 1)));
    CHECK_CUDA(cudaMalloc((void**)&d_out, row_bytes * (n_rows > 0 ? n_rows : 1)));
    cudaStream_t st;
    CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int img = 0; img < c->images; ++img) { /* every rank, the same sequence: each restore contains the cross-rank barriers */
        CHECK_FDR(fdr_synth_rows_device_u8(d_in, 0xF17E0004u, img, 3, c->rows, c->cols, first_row, n_rows, st));
        CHECK_FDR(fdr_shard_restore_rows(sh, d_in, d_out, st));
    }
    int timed_out = 0;
    CHECK_FDR(fdr_shard_sync_status(sh, st, &timed_out)); /* synchronises st */
    if (timed_out) {
        fprintf(stderr, "Error: rank %d: a cross-rank barrier timed out\n", rank);
        _exit(4);
    }
    if (n_rows > 0) { /* every rank writes its own rows: no gather to a root (the reference gathers, fft_mpi.cpp:455-463) */
        unsigned char* h = (unsigned char*)malloc(row_bytes * n_rows);
        CHECK_FDR(fdr_memcpy(h, d_out, row_bytes * n_rows, 1));
        int fd = open(c->out_path, O_WRONLY | O_CREAT, 0644);
        if (fd < 0 || pwrite(fd, h, row_bytes * n_rows, (off_t)(row_bytes * first_row)) != (ssize_t)(row_bytes * n_rows)) {
            fprintf(stderr, "Error: rank %d cannot write %s\n", rank, c->out_path);
            _exit(5);
        }
        close(fd);
        free(h);
    }
    /* tear-down: unmap the peers' slabs everywhere BEFORE any rank frees its own */
    for (int r = 0; r < world; ++r)
        if (r != rank) CHECK_FDR(fdr_ipc_close(peers[r]));
    MPI_Barrier(MPI_COMM_WORLD);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaStreamDestroy(st);
    CHECK_FDR(fdr_shard_destroy(sh));
}

int main(int argc, char** argv) {
    if (argc != 8) {
        fprintf(stderr, "Usage: %s <ranks> <rows> <cols> <psf-length> <psf-angle> <images> <out.bin>\n", argv[0]);
        return 2;
    }
    struct cfg c;
    const int ranks = atoi(argv[1]);
    c.rows = atoi(argv[2]);
    c.cols = atoi(argv[3]);
    c.psf_len = atoi(argv[4]);
    c.psf_angle = atof(argv[5]);
    c.images = atoi(argv[6]);
    c.out_path = argv[7];
    if (ranks < 1 || ranks > MAXRANKS || c.rows < 1 || c.cols < 1 || c.images < 1) return 2;
    /* CUDA must not be initialised before the ranks are forked: nothing above touches the device */
    return mpi_standin_launch(ranks, 1 << 16, rank_main, &c) == 0 ? 0 : 1;
}
