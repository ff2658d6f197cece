// Microbenchmark: issue rate of scalar FFMA / FADD vs packed FFMA2 / FADD2 / FMUL2 on sm_100a,
// alone and mixed with ALU (LOP3) work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float2* out, int iters, float2 s) {
    float2 a[8];
    unsigned int z[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, s.x, s.y); a[i].y = fmaf(a[i].y, s.x, s.y); }          // 2 FFMA
                if (MODE == 1) a[i] = __ffma2_rn(a[i], s, s);                                                 // 1 FFMA2
                if (MODE == 2) { a[i].x = a[i].x + s.x; a[i].y = a[i].y + s.y; }                              // 2 FADD
                if (MODE == 3) a[i] = __fadd2_rn(a[i], s);                                                    // 1 FADD2
                if (MODE == 4) { a[i] = __ffma2_rn(a[i], s, s); z[i & 3] = (z[i & 3] ^ (unsigned)it) + 0x9e37u; }  // FFMA2 + 2 ALU-ish
                if (MODE == 5) { a[i].x = fmaf(a[i].x, s.x, s.y); a[i].y = fmaf(a[i].y, s.x, s.y); z[i & 3] = (z[i & 3] ^ (unsigned)it) + 0x9e37u; }
            }
        }
    }
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc.x += a[i].x; acc.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(acc.x + z[0] + z[1], acc.y + z[2] + z[3]);
}
template <int MODE> void run(const char* name, float2* d) {
    const int iters = 2000, blocks = 148 * 4, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 10, make_float2(1.0001f, 0.5f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, make_float2(1.0001f, 0.5f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads / 32 * iters * 64;  // "complex ops" (pairs) per warp
    printf("%-28s %.3f ms  %.2f pair-ops/clk/SM (at 1.9 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 4 * 512 * sizeof(float2));
    run<0>("2x FFMA", d); run<1>("FFMA2", d); run<2>("2x FADD", d); run<3>("FADD2", d); run<4>("FFMA2 + int ops", d); run<5>("2x FFMA + int ops", d);
    return 0;
}
