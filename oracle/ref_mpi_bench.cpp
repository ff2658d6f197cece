// TEST / BENCH INFRASTRUCTURE ONLY -- not part of the product path.
//
// Runs the reference's MPI backend (fft/fft_mpi.cpp, compiled unmodified against oracle/mpi_standin/mpi.h)
// as P forked ranks on this host, following the reference driver's loop (mpi.cpp:95-111): rank 0 pads each
// channel and calls fft_mpi::wienerDeblur_myfft, the workers call it with empty Mats.
//   ref_mpi_bench <nprocs> <planes.f32> <n_planes> <rows> <cols> <psf.f32> <psf_size> <K> <out.f32>
// planes are already padded to powers of two.  Prints "mpi_ms <wall ms of the loop on rank 0>".
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft/fft.hpp"
#include "mpi.h"

struct Job {
    int n_planes, rows, cols, psf_size;
    float K;
    std::vector<float> planes, psf, out;
    double ms;
    const char* out_path;
};

static void body(void* arg) {
    Job* j = static_cast<Job*>(arg);
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Barrier(MPI_COMM_WORLD);
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < j->n_planes; ++i) {
        if (rank == 0) {
            cv::Mat img(j->rows, j->cols, CV_32F), psf(j->psf_size, j->psf_size, CV_32F);
            memcpy(img.data, j->planes.data() + (size_t)i * j->rows * j->cols, sizeof(float) * (size_t)j->rows * j->cols);
            memcpy(psf.data, j->psf.data(), sizeof(float) * j->psf.size());
            cv::Mat r = fft_mpi::wienerDeblur_myfft(img, psf, j->K);
            memcpy(j->out.data() + (size_t)i * j->rows * j->cols, r.data, sizeof(float) * (size_t)j->rows * j->cols);
        } else {
            fft_mpi::wienerDeblur_myfft(cv::Mat(), cv::Mat(), j->K);
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (rank == 0) {
        j->ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        FILE* f = fopen(j->out_path, "wb");
        if (f) {
            fwrite(j->out.data(), sizeof(float), j->out.size(), f);
            fclose(f);
        }
        printf("mpi_ms %.3f\n", j->ms);
    }
}

int main(int argc, char** argv) {
    if (argc != 10) {
        fprintf(stderr, "usage: ref_mpi_bench <nprocs> <planes.f32> <n_planes> <rows> <cols> <psf.f32> <psf_size> <K> <out.f32>\n");
        return 2;
    }
    Job j;
    const int nprocs = atoi(argv[1]);
    j.n_planes = atoi(argv[3]);
    j.rows = atoi(argv[4]);
    j.cols = atoi(argv[5]);
    j.psf_size = atoi(argv[7]);
    j.K = (float)atof(argv[8]);
    j.out_path = argv[9];
    j.planes.resize((size_t)j.n_planes * j.rows * j.cols);
    j.psf.resize((size_t)j.psf_size * j.psf_size);
    j.out.resize(j.planes.size());
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(j.planes.data(), sizeof(float), j.planes.size(), f) != j.planes.size()) return 3;
    fclose(f);
    f = fopen(argv[6], "rb");
    if (!f || fread(j.psf.data(), sizeof(float), j.psf.size(), f) != j.psf.size()) return 3;
    fclose(f);
    // largest collective: Alltoallv / Scatterv / Gatherv of one complex plane
    const size_t arena = (size_t)j.rows * j.cols * 2 * sizeof(float) + (1 << 20);
    return mpi_standin_launch(nprocs, arena, body, &j) == 0 ? 0 : 1;
}
