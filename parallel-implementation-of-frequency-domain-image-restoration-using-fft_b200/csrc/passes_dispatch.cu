// passes_dispatch.cu -- run-time length -> instantiated kernel.
#include "passes.h"

namespace fdr {
#define X(LOGN)                                                              \
    cudaError_t launch_row_pass_##LOGN(const RowPassArgs&, cudaStream_t);    \
    cudaError_t launch_col_pass_##LOGN(const ColPassArgs&, cudaStream_t);    \
    cudaError_t configure_pass_##LOGN();
#define FDR_ALL_LOGNS X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)
FDR_ALL_LOGNS
#undef X

static int ilog2_exact(int n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

cudaError_t launch_row_pass(const RowPassArgs& a, cudaStream_t s) {
    switch (ilog2_exact(a.n)) {
#define X(LOGN) \
    case LOGN: return launch_row_pass_##LOGN(a, s);
        FDR_ALL_LOGNS
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_col_pass(const ColPassArgs& a, cudaStream_t s) {
    switch (ilog2_exact(a.n)) {
#define X(LOGN) \
    case LOGN: return launch_col_pass_##LOGN(a, s);
        FDR_ALL_LOGNS
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t configure_pass_kernels() {
    cudaError_t e;
#define X(LOGN)                  \
    e = configure_pass_##LOGN(); \
    if (e != cudaSuccess) return e;
    FDR_ALL_LOGNS
#undef X
    return cudaSuccess;
}

int col_pass_tile_width(int n) {
    int l = ilog2_exact(n);
    return l <= 6 ? 32 : l <= 8 ? 16 : l <= 10 ? 8 : l == 11 ? 4 : l == 12 ? 2 : 1;
}
}  // namespace fdr
