// capi.cu -- plan management and the extern "C" surface declared in include/fdr_b200.h.
//
// Host-side orchestration only: which fused pass runs on which chunk of planes, staging for the
// host entry points, error translation.  No arithmetic on pixel data happens on the CPU; there
// is no CPU fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "capi_internal.h"

using namespace fdr;

namespace fdr {
static thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
const char* last_error_text() { return g_last_error.c_str(); }
}  // namespace fdr

namespace {

int ensure_device(int device) {
    FDR_CUDA(cudaSetDevice(device));
    return FDR_OK;
}

}  // namespace

struct fdr_plan {
    int device = 0;
    int H = 0, W = 0, C = 0, max_images = 0;
    int Rp = 0, Cp = 0;
    int chunk_images_user = 0;
    float K = 0.f;
    int psf_rows = 0, psf_cols = 0;
    bool have_wiener = false;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the pipelined host path
    cudaEvent_t ev_in[2] = {}, ev_cmp[2] = {}, ev_out[2] = {};
    cudaEvent_t ev[8] = {};
    // device workspace
    DevBuf<float2> spec;          // chunk pairs x Rp x Cp
    DevBuf<float> raw;            // chunk units x H x W
    DevBuf<unsigned int> mm;      // chunk units x FDR_MINMAX_SLOTS x 2
    DevBuf<float2> ss;            // chunk units
    DevBuf<float> mmf;            // chunk units x 2
    DevBuf<double> wb_sums;       // chunk images x 2 (Lab white balance)
    int white_balance = 0;        // 8-bit outputs go through the Lab white-balance stage (gpu.cpp:123-134)
    DevBuf<float2> wiener;        // Rp x Cp (digit-swapped row order when col_split is set)
    DevBuf<float2> wiener_tiled;  // tile-major copy for the wide column kernel (Rp == 2048), passes.h
    DevBuf<float2> wiener_nat;    // natural-order copy, built lazily for the parity-gate API of long-column plans
    bool col_split = false;       // long columns: four-step column pass (col_split.cuh)
    // Half-plane path for the odd colour plane of a chunk (passes.h, ROW_*_HALF): its own Wiener factors, built by the same
    // kernels at half pitch (so their row order is whatever the column launcher of that geometry expects).
    bool half_ok = false;         // geometry allows it and FDR_HALF != 0
    bool col_split_half = false;
    DevBuf<float2> wiener_half;        // Rp x Cp/2: columns 0 .. Cp/2-1
    DevBuf<float2> wiener_half_tiled;  // tile-major copy when the wide column kernel serves pitch Cp/2
    DevBuf<float2> wiener_nyq;         // Rp: column Cp/2, natural row order (plain column kernel)
    cudaStream_t s_half = nullptr;     // the lone plane's passes run beside the pairs' passes
    cudaEvent_t ev_half_fork = nullptr, ev_half_join = nullptr;
    // one workspace per plan: calls on different streams are ordered through this event (include/fdr_b200.h, "Streams")
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    DevBuf<float> psf;            // psf_rows x psf_cols
    int lanes = 1;                    // chunks in flight on separate streams (FDR_LANES=1..4); with large chunks one is best
    cudaStream_t lane_stream[4] = {};
    cudaEvent_t lane_fork = nullptr, lane_join[4] = {};
    int last_lane = 0;
    int ws_units = 0;                 // per-lane workspace capacity in units
    const float2* tw_rows = nullptr;  // twiddles for length Cp (row passes)
    const float2* tw_cols = nullptr;  // twiddles for length Rp (column passes)
    // staging for the host entry points
    DevBuf<uint8_t> d_in_u8, d_out_u8;
    DevBuf<float> d_in_f32, d_out_f32;
    PinBuf<uint8_t> h_u8_in, h_u8_out;
    PinBuf<float> h_f32_in, h_f32_out;
    float profile_ms[6] = {0, 0, 0, 0, 0, 0};
    long long launches = 0;
    int last_units = 0;
    // optional per-kernel timing (event pair around every pass launch)
    bool ktiming = false;
    struct KRec {
        int kind;
        cudaEvent_t a, b;
        double bytes;
    };
    std::vector<KRec> krecs;
    std::vector<cudaEvent_t> ev_pool;

    size_t plane_elems() const { return (size_t)Rp * Cp; }
    // Images per device chunk.  Measured on B200 (profiles/r1b/chunk_sweep_small.txt): the passes are bound by the
    // SM-side load/store pipe or by HBM, not helped by L2 residency between passes, and every launch pays
    // a partial last wave -- so large chunks win: 32 images of 2048^2 (1.5 GB of spectrum) run 12 % faster
    // than the 96 MB chunks that fit the L2.  An even unit count per chunk keeps plane pairs inside a chunk.
    int chunk_images(int n_images_total) const {
        int ci = chunk_images_user;
        if (ci <= 0) {
            const double target = 1536.0 * 1024 * 1024;
            const double per_image = 0.5 * C * (double)plane_elems() * sizeof(float2);
            ci = (int)(target / per_image);
            if (ci < 1) ci = 1;
            const char* env = getenv("FDR_CHUNK_IMAGES");
            if (env && atoi(env) > 0) ci = atoi(env);
        }
        if ((ci * C) % 2 && ci < n_images_total) ci += 1;
        if (ci > n_images_total) ci = n_images_total;
        ci = cap_grid(ci);
        return ci < 1 ? 1 : ci;
    }
    // plane pairs, planes and images of a chunk index gridDim.y of the pass kernels: stay below 65535 (huge batches of tiny images)
    int cap_grid(int ci) const {
        const int max_images = 65534 / (C > 0 ? C : 1);
        if (ci > max_images) {
            ci = max_images;
            if ((ci * C) % 2) ci -= 1;
        }
        return ci;
    }
    // Images per chunk of the pipelined HOST entry point: that path is PCIe-bound, so small chunks
    // (about 75 MB of 8-bit input: 6 images of 2048^2) keep the fill and drain of the
    // H2D -> compute -> D2H pipeline short; measured best among 2..32.
    int host_chunk_images(int n_images_total) const {
        int ci = chunk_images_user;
        if (ci <= 0) {
            const double per_image = (double)H * W * C;
            ci = (int)(75.0e6 / per_image);
            if (ci < 1) ci = 1;
            const char* env = getenv("FDR_HOST_CHUNK_IMAGES");
            if (env && atoi(env) > 0) ci = atoi(env);
        }
        if ((ci * C) % 2 && ci < n_images_total) ci += 1;
        if (ci > n_images_total) ci = n_images_total;
        ci = cap_grid(ci);
        return ci < 1 ? 1 : ci;
    }
};

namespace {

struct InputDesc {
    int mode;  // ROW_IN_PAIR_F32 / ROW_IN_PAIR_U8
    const float* f32 = nullptr;
    long long unit_stride = 0, row_stride = 0;
    const uint8_t* u8 = nullptr;
};

struct KernelTimer {  // records an event pair around one launch when kernel timing is on
    fdr_plan* p;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    int kind;
    double bytes;
    static cudaEvent_t get(fdr_plan* p) {
        if (!p->ev_pool.empty()) {
            cudaEvent_t e = p->ev_pool.back();
            p->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    KernelTimer(fdr_plan* p_, cudaStream_t s_, int kind_, double bytes_) : p(p_), s(s_), kind(kind_), bytes(bytes_) {
        if (p->ktiming) {
            a = get(p);
            b = get(p);
            cudaEventRecord(a, s);
        }
    }
    ~KernelTimer() {
        if (p->ktiming) {
            cudaEventRecord(b, s);
            p->krecs.push_back({kind, a, b, bytes});
        }
    }
};

int ensure_workspace(fdr_plan* p, int chunk_units) {
    if (chunk_units < p->ws_units) chunk_units = p->ws_units;  // lane offsets must stay valid: capacity only grows
    const int pairs = (chunk_units + 1) / 2;
    const size_t L = (size_t)p->lanes;
    FDR_TRY(p->spec.ensure(L * pairs * p->plane_elems()));
    FDR_TRY(p->raw.ensure(L * chunk_units * p->H * p->W));
    FDR_TRY(p->mm.ensure(L * chunk_units * 2 * FDR_MINMAX_SLOTS));
    FDR_TRY(p->ss.ensure(L * chunk_units));
    FDR_TRY(p->mmf.ensure(L * chunk_units * 2));
    FDR_TRY(p->wb_sums.ensure(L * chunk_units * 2));
    p->ws_units = chunk_units;
    if (p->lanes > 1 && !p->lane_fork) {
        FDR_CUDA(cudaEventCreateWithFlags(&p->lane_fork, cudaEventDisableTiming));
        for (int i = 0; i < p->lanes; ++i) {
            FDR_CUDA(cudaStreamCreateWithFlags(&p->lane_stream[i], cudaStreamNonBlocking));
            FDR_CUDA(cudaEventCreateWithFlags(&p->lane_join[i], cudaEventDisableTiming));
        }
    }
    return FDR_OK;
}

// Restores units [0, n_units) (unit = one colour plane; image i = units i*C .. i*C+C-1).
// out_f32: normalised planes [unit][H][W] or NULL; out_u8: interleaved images or NULL.
int restore_units_device(fdr_plan* p, const InputDesc& in, float* out_f32, uint8_t* out_u8, long long n_units,
                         cudaStream_t s) {
    if (!p->have_wiener) return set_error(FDR_E_STATE, "no PSF set: call fdr_plan_set_psf_* first");
    if (n_units <= 0) return FDR_OK;
    const int C = p->C;
    if (out_u8 && (n_units % C)) return set_error(FDR_E_INVALID, "8-bit output needs whole images (%lld planes, %d channels)", n_units, C);
    const long long n_images = (n_units + C - 1) / C;
    long long chunk_units = (long long)p->chunk_images((int)n_images) * C;
    if (chunk_units > n_units) chunk_units = n_units;
    if (!out_u8 && (chunk_units % 2) && chunk_units < n_units) chunk_units += 1;
    FDR_TRY(ensure_workspace(p, (int)chunk_units));
    p->launches = 0;
    if (p->have_last && p->last_stream != s) FDR_CUDA(cudaStreamWaitEvent(s, p->ev_last, 0));  // the workspace is shared
    const long long HW = (long long)p->H * p->W;
    const int L = (n_units > chunk_units) ? p->lanes : 1;  // several chunks in flight on separate streams
    cudaStream_t s_caller = s;
    if (L > 1) {
        FDR_CUDA(cudaEventRecord(p->lane_fork, s_caller));
        for (int i = 0; i < L; ++i) FDR_CUDA(cudaStreamWaitEvent(p->lane_stream[i], p->lane_fork, 0));
    }
    const size_t ws_pairs = (size_t)(p->ws_units + 1) / 2;
    long long chunk_index = 0;
    for (long long base = 0; base < n_units; base += chunk_units, ++chunk_index) {
        const int nu = (int)((n_units - base < chunk_units) ? (n_units - base) : chunk_units);
        const int np = (nu + 1) / 2;
        const int lane = (int)(chunk_index % L);
        if (L > 1) s = p->lane_stream[lane];
        float2* const spec_l = p->spec.p + (size_t)lane * ws_pairs * p->plane_elems();
        float* const raw_l = p->raw.p + (size_t)lane * p->ws_units * HW;
        unsigned int* const mm_l = p->mm.p + (size_t)lane * p->ws_units * 2 * FDR_MINMAX_SLOTS;
        float2* const ss_l = p->ss.p + (size_t)lane * p->ws_units;
        float* const mmf_l = p->mmf.p + (size_t)lane * p->ws_units * 2;
        p->last_lane = lane;
        // the slots are re-armed by CTA 0 of every pair in pass 1 (RowPassArgs::mm_reset): no separate launch
        // An odd plane count leaves one plane without a partner: it takes the half-plane path (passes.h) on a side stream,
        // beside the pairs, instead of travelling as a half-empty complex pair.
        const bool lone = (nu & 1) && p->half_ok;
        const int npf = lone ? nu / 2 : np;  // full pairs
        const double px_in1 = (double)p->H * p->W * (in.mode == ROW_IN_PAIR_U8 ? 1.0 : 4.0);
        const double P = (double)p->plane_elems();
        cudaStream_t sh = s;
        if (lone && npf > 0) {
            sh = p->s_half;
            FDR_CUDA(cudaEventRecord(p->ev_half_fork, s));
            FDR_CUDA(cudaStreamWaitEvent(sh, p->ev_half_fork, 0));
        }
        if (npf > 0) {
            RowPassArgs r1{};
            r1.n = p->Cp;
            r1.nrows = p->H;
            r1.npairs = npf;
            r1.in_mode = in.mode;
            r1.out_mode = ROW_OUT_COMPLEX;
            r1.in_f32 = in.f32;
            r1.in_unit_stride = in.unit_stride;
            r1.in_row_stride = in.row_stride;
            r1.in_u8 = in.u8;
            r1.channels = C;
            r1.img_rows = p->H;
            r1.img_cols = p->W;
            r1.unit_base = base;
            r1.units_total = n_units;
            r1.cout = spec_l;
            r1.cplane = (long long)p->plane_elems();
            r1.tw = p->tw_rows;
            r1.mm_reset = mm_l;
            r1.local_units = nu;
            {
                KernelTimer kt(p, s, 0, px_in1 * (lone ? nu - 1 : nu) + 8.0 * p->H * p->Cp * npf);
                FDR_CUDA(launch_row_pass(r1, s));
            }

            ColPassArgs c2{};
            c2.n = p->Rp;
            c2.pitch = p->Cp;
            c2.npairs = npf;
            c2.mode = COL_WIENER;
            c2.rows_valid = p->H;
            c2.data = spec_l;
            c2.cplane = (long long)p->plane_elems();
            c2.wiener = p->wiener.p;
            c2.wiener_tiled = p->wiener_tiled.p;
            c2.K = p->K;
            c2.tw = p->tw_cols;
            {
                KernelTimer kt(p, s, 1, (8.0 * p->H * p->Cp + 16.0 * P) * npf);
                if (p->col_split) {
                    int nl = 0;
                    FDR_CUDA(launch_col_split(c2, s, &nl));
                    p->launches += nl - 1;
                } else {
                    FDR_CUDA(launch_col_pass(c2, s));
                }
            }

            RowPassArgs r3{};
            r3.n = p->Cp;
            r3.nrows = p->Rp;
            r3.npairs = npf;
            r3.in_mode = ROW_IN_COMPLEX;
            r3.out_mode = ROW_OUT_REAL_PAIR;
            r3.cin = spec_l;
            r3.cplane = (long long)p->plane_elems();
            r3.unit_base = base;
            r3.units_total = n_units;
            r3.raw = raw_l;
            r3.raw_unit_stride = HW;
            r3.raw_rows = p->H;
            r3.raw_cols = p->W;
            r3.minmax = mm_l;
            r3.local_units = nu;
            r3.tw = p->tw_rows;
            {
                KernelTimer kt(p, s, 2, 8.0 * P * npf + 4.0 * HW * (lone ? nu - 1 : nu));
                FDR_CUDA(launch_row_pass(r3, s));
            }
            p->launches += 3;
        }
        if (lone) {
            // workspace of the lone plane: the pair slot after the full pairs holds [Rp][Cp/2] + the Nyquist column [Rp]
            float2* const hbuf = spec_l + (size_t)npf * p->plane_elems();
            float2* const nyq = hbuf + (size_t)p->Rp * (p->Cp / 2);
            const int D1 = (p->H + 1) / 2;
            RowPassArgs h1{};
            h1.n = p->Cp;
            h1.nrows = D1;
            h1.npairs = 1;
            h1.pair_base = nu - 1;  // half-plane forms index planes
            h1.in_mode = (in.mode == ROW_IN_PAIR_U8) ? ROW_IN_ROWS2_U8 : ROW_IN_ROWS2_F32;
            h1.out_mode = ROW_OUT_HALF;
            h1.in_f32 = in.f32;
            h1.in_unit_stride = in.unit_stride;
            h1.in_row_stride = in.row_stride;
            h1.in_u8 = in.u8;
            h1.channels = C;
            h1.img_rows = p->H;
            h1.img_cols = p->W;
            h1.unit_base = base;
            h1.units_total = n_units;
            h1.tw = p->tw_rows;
            h1.mm_reset = mm_l;
            h1.local_units = nu;
            h1.pair_dist = D1;
            h1.rows_in = p->H;
            h1.hp_rows_store = p->H;
            h1.hp_peers[0] = hbuf;
            h1.hp_shift = ilog2(p->Cp / 2);
            h1.hp_plane = 0;
            h1.nyq_peers[0] = nyq;
            h1.nyq_world = 1;
            h1.nyq_plane = 0;
            {
                KernelTimer kt(p, sh, 0, px_in1 + 4.0 * p->H * p->Cp);
                FDR_CUDA(launch_row_pass(h1, sh));
            }
            ColPassArgs ch{};
            ch.n = p->Rp;
            ch.pitch = p->Cp / 2;
            ch.npairs = 1;
            ch.mode = COL_WIENER;
            ch.rows_valid = p->H;
            ch.data = hbuf;
            ch.cplane = (long long)p->Rp * (p->Cp / 2);
            ch.wiener = p->wiener_half.p;
            ch.wiener_tiled = p->wiener_half_tiled.p;
            ch.K = p->K;
            ch.tw = p->tw_cols;
            {
                KernelTimer kt(p, sh, 1, 4.0 * p->H * p->Cp + 8.0 * P);
                if (p->col_split_half) {
                    int nl = 0;
                    FDR_CUDA(launch_col_split(ch, sh, &nl));
                    p->launches += nl - 1;
                } else {
                    FDR_CUDA(launch_col_pass(ch, sh));
                }
                ColPassArgs cn = ch;  // the Nyquist column: one more column of the same problem
                cn.pitch = 1;
                cn.data = nyq;
                cn.cplane = p->Rp;
                cn.wiener = p->wiener_nyq.p;
                cn.wiener_tiled = nullptr;
                FDR_CUDA(launch_col_pass(cn, sh));
            }
            RowPassArgs h3{};
            h3.n = p->Cp;
            h3.nrows = p->Rp / 2;
            h3.npairs = 1;
            h3.pair_base = nu - 1;
            h3.in_mode = ROW_IN_HALF;
            h3.out_mode = ROW_OUT_REAL_ROWS2;
            h3.unit_base = base;
            h3.units_total = n_units;
            h3.raw = raw_l;
            h3.raw_unit_stride = HW;
            h3.raw_rows = p->H;
            h3.raw_cols = p->W;
            h3.minmax = mm_l;
            h3.local_units = nu;
            h3.tw = p->tw_rows;
            h3.pair_dist = p->Rp / 2;
            h3.hp_peers[0] = hbuf;
            h3.hp_shift = ilog2(p->Cp / 2);
            h3.hp_plane = 0;
            h3.nyq_peers[0] = nyq;
            h3.nyq_world = 1;
            h3.nyq_plane = 0;
            {
                KernelTimer kt(p, sh, 2, 4.0 * P + 4.0 * HW);
                FDR_CUDA(launch_row_pass(h3, sh));
            }
            p->launches += 4;
            if (sh != s) {
                FDR_CUDA(cudaEventRecord(p->ev_half_join, sh));
                FDR_CUDA(cudaStreamWaitEvent(s, p->ev_half_join, 0));
            }
        }

        // a chunk of one or two BGR images: the pack folds the slots itself (one launch less on the launch-bound small images)
        const bool fused_pack = out_u8 && !out_f32 && !(p->white_balance && C == 3) &&
                                pack_u8_c3_fused_applicable(raw_l, HW, out_u8 + base * HW, nu / C, C, p->H, p->W);
        if (!fused_pack) {
            FDR_CUDA(launch_minmax_finalize(mm_l, ss_l, mmf_l, nu, s));
            p->launches += 1;
        }
        if (fused_pack) {
            KernelTimer kt(p, s, 3, 5.0 * HW * nu);
            FDR_CUDA(launch_pack_u8_c3_fused(raw_l, HW, mm_l, ss_l, mmf_l, out_u8 + base * HW, nu / C, p->H, p->W, s));
            p->launches += 1;
        } else if (out_u8 && p->white_balance && C == 3) {
            KernelTimer kt(p, s, 3, (2 * 12.0 + 3.0 + (in.mode == ROW_IN_PAIR_U8 ? 3.0 : 12.0)) * HW * (nu / 3));
            const uint8_t* o8 = in.mode == ROW_IN_PAIR_U8 ? in.u8 + base * HW : nullptr;
            const float* of = in.mode == ROW_IN_PAIR_U8 ? nullptr : in.f32 + base * in.unit_stride;
            FDR_CUDA(launch_white_balance_pack_u8(raw_l, HW, ss_l, o8, of, in.unit_stride, p->wb_sums.p + (size_t)lane * p->ws_units * 2,
                                                  out_u8 + base * HW, nu / 3, p->H, p->W, s));
            p->launches += 3;
        } else if (out_u8) {
            KernelTimer kt(p, s, 3, 5.0 * HW * nu);
            FDR_CUDA(launch_pack_u8(raw_l, HW, ss_l, out_u8 + base * HW, nu / C, C, p->H, p->W, s));
            p->launches += 1;
        }
        if (out_f32) {
            KernelTimer kt(p, s, 3, 8.0 * HW * nu);
            FDR_CUDA(launch_normalize_f32(raw_l, HW, ss_l, out_f32 + base * HW, HW, nu, p->H, p->W, s));
            p->launches += 1;
        }
        p->last_units = nu;
    }
    if (L > 1) {
        for (int i = 0; i < L; ++i) {
            FDR_CUDA(cudaEventRecord(p->lane_join[i], p->lane_stream[i]));
            FDR_CUDA(cudaStreamWaitEvent(s_caller, p->lane_join[i], 0));
        }
    }
    FDR_CUDA(cudaEventRecord(p->ev_last, s_caller));
    p->last_stream = s_caller;
    p->have_last = true;
    return FDR_OK;
}

int build_wiener_into(fdr_plan* p, DevBuf<float2>& dst, bool split) {
    if (p->psf_rows > p->Rp || p->psf_cols > p->Cp)
        return set_error(FDR_E_INVALID, "PSF %dx%d larger than the padded image %dx%d", p->psf_rows, p->psf_cols, p->Rp, p->Cp);
    FDR_TRY(dst.ensure(p->plane_elems()));
    FDR_TRY(p->spec.ensure(p->plane_elems()));
    cudaStream_t s = p->stream;
    if (p->have_last && p->last_stream != s) FDR_CUDA(cudaStreamWaitEvent(s, p->ev_last, 0));  // restores still using the workspace
    RowPassArgs r{};
    r.n = p->Cp;
    r.nrows = p->psf_rows;
    r.npairs = 1;
    r.in_mode = ROW_IN_PAIR_F32;
    r.out_mode = ROW_OUT_COMPLEX;
    r.in_f32 = p->psf.p;
    r.in_unit_stride = (long long)p->psf_rows * p->psf_cols;
    r.in_row_stride = p->psf_cols;
    r.channels = 1;
    r.img_rows = p->psf_rows;
    r.img_cols = p->psf_cols;
    r.unit_base = 0;
    r.units_total = 1;
    r.cout = p->spec.p;
    r.cplane = (long long)p->plane_elems();
    r.tw = p->tw_rows;
    FDR_CUDA(launch_row_pass(r, s));
    ColPassArgs c{};
    c.n = p->Rp;
    c.pitch = p->Cp;
    c.npairs = 1;
    c.mode = COL_MAKE_WIENER;
    c.rows_valid = p->psf_rows;
    c.data = p->spec.p;
    c.cplane = (long long)p->plane_elems();
    c.wiener_out = dst.p;
    c.K = p->K;
    c.tw = p->tw_cols;
    if (split) {
        FDR_CUDA(launch_col_split(c, s, nullptr));
    } else {
        FDR_CUDA(launch_col_pass(c, s));
    }
    FDR_CUDA(cudaStreamSynchronize(s));
    return FDR_OK;
}

struct ScopedTimer {  // event pair on a stream, accumulating into a bucket (Profiler, fft_gpu.cu:17-57)
    cudaEvent_t a, b;
    cudaStream_t s;
    float* dst;
    ScopedTimer(fdr_plan* p, int slot, cudaStream_t st, float* d) : a(p->ev[2 * slot]), b(p->ev[2 * slot + 1]), s(st), dst(d) {
        cudaEventRecord(a, s);
    }
    void stop() {
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        *dst += ms;
    }
};

// Wiener factors of the half-plane path: columns 0 .. Cp/2-1 as a plane of pitch Cp/2 and the Nyquist column Cp/2, built
// by the same kernels at that pitch (fft_serial.cpp:166-224 on those columns only).
int build_wiener_half(fdr_plan* p) {
    const int Ch = p->Cp / 2;
    const size_t half_elems = (size_t)p->Rp * Ch;
    cudaStream_t s = p->stream;
    FDR_TRY(p->wiener_half.ensure(half_elems));
    FDR_TRY(p->wiener_nyq.ensure((size_t)p->Rp));
    DevBuf<float2> tmp;  // scratch of the in-place column passes: [Rp][Cp/2] + [Rp]
    FDR_TRY(tmp.ensure(half_elems + p->Rp));
    int rc = FDR_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == FDR_OK) rc = set_error(FDR_E_CUDA, "half-plane Wiener build (%s): %s", what, cudaGetErrorString(e));
    };
    RowPassArgs r{};  // the PSF's row spectrum again (the full build consumed it)
    r.n = p->Cp;
    r.nrows = p->psf_rows;
    r.npairs = 1;
    r.in_mode = ROW_IN_PAIR_F32;
    r.out_mode = ROW_OUT_COMPLEX;
    r.in_f32 = p->psf.p;
    r.in_unit_stride = (long long)p->psf_rows * p->psf_cols;
    r.in_row_stride = p->psf_cols;
    r.channels = 1;
    r.img_rows = p->psf_rows;
    r.img_cols = p->psf_cols;
    r.units_total = 1;
    r.cout = p->spec.p;
    r.cplane = (long long)p->plane_elems();
    r.tw = p->tw_rows;
    cu(launch_row_pass(r, s), "rows");
    cu(cudaMemcpy2DAsync(tmp.p, (size_t)Ch * sizeof(float2), p->spec.p, (size_t)p->Cp * sizeof(float2), (size_t)Ch * sizeof(float2),
                         p->psf_rows, cudaMemcpyDeviceToDevice, s), "left half");
    cu(cudaMemcpy2DAsync(tmp.p + half_elems, sizeof(float2), p->spec.p + Ch, (size_t)p->Cp * sizeof(float2), sizeof(float2), p->psf_rows,
                         cudaMemcpyDeviceToDevice, s), "nyquist column");
    ColPassArgs c{};
    c.n = p->Rp;
    c.pitch = Ch;
    c.npairs = 1;
    c.mode = COL_MAKE_WIENER;
    c.rows_valid = p->psf_rows;
    c.data = tmp.p;
    c.cplane = (long long)half_elems;
    c.wiener_out = p->wiener_half.p;
    c.K = p->K;
    c.tw = p->tw_cols;
    if (rc == FDR_OK) cu(p->col_split_half ? launch_col_split(c, s, nullptr) : launch_col_pass(c, s), "columns");
    ColPassArgs n = c;
    n.pitch = 1;
    n.data = tmp.p + half_elems;
    n.cplane = p->Rp;
    n.wiener_out = p->wiener_nyq.p;
    if (rc == FDR_OK) cu(launch_col_pass(n, s), "nyquist");
    p->wiener_half_tiled.release();
    ColPassArgs probe{};
    probe.n = p->Rp;
    probe.pitch = Ch;
    probe.mode = COL_WIENER;
    if (rc == FDR_OK && !p->col_split_half && col_wide_applicable(probe)) {
        rc = p->wiener_half_tiled.ensure(half_elems);
        if (rc == FDR_OK) cu(launch_wiener_retile(p->wiener_half.p, p->wiener_half_tiled.p, p->Rp, Ch, s), "retile");
    }
    cu(cudaStreamSynchronize(s), "sync");
    tmp.release();
    return rc;
}

int build_wiener(fdr_plan* p) {
    p->wiener_nat.release();
    FDR_TRY(build_wiener_into(p, p->wiener, p->col_split));
    p->wiener_tiled.release();
    {
        ColPassArgs probe{};
        probe.n = p->Rp;
        probe.pitch = p->Cp;
        probe.mode = COL_WIENER;
        if (!p->col_split && col_wide_applicable(probe)) {
            FDR_TRY(p->wiener_tiled.ensure(p->plane_elems()));
            FDR_CUDA(launch_wiener_retile(p->wiener.p, p->wiener_tiled.p, p->Rp, p->Cp, p->stream));
            FDR_CUDA(cudaStreamSynchronize(p->stream));
        }
    }
    if (p->half_ok) FDR_TRY(build_wiener_half(p));
    p->have_wiener = true;
    return FDR_OK;
}

}  // namespace

// Inverse rotation matrix exactly as OpenCV derives it (getRotationMatrix2D + warpAffine's
// inversion, both in double); utils.hpp:16-22.
PsfAffine fdr::motion_affine(int size, double angle_deg) {
    const double CVPI = 3.1415926535897932384626433832795;
    const double cx = (double)(float)(size / 2), cy = (double)(float)(size / 2);
    double ang = angle_deg * (CVPI / 180);
    double alpha = std::cos(ang), beta = std::sin(ang);
    double M[6] = {alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy};
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11;
    M[1] *= -D;
    M[3] *= -D;
    M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5];
    double b2 = -M[3] * M[2] - M[4] * M[5];
    PsfAffine a;
    a.a00 = M[0];
    a.a01 = M[1];
    a.b0 = b1;
    a.a10 = M[3];
    a.a11 = M[4];
    a.b1 = b2;
    return a;
}


extern "C" {

__attribute__((visibility("default"))) const char* fdr_last_error(void) { return fdr::last_error_text(); }
__attribute__((visibility("default"))) int fdr_version(void) { return 100; }

__attribute__((visibility("default"))) int fdr_device_count(int* count) {
    if (!count) return set_error(FDR_E_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return set_error(FDR_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return set_error(FDR_E_INVALID, "ptr is NULL");
    FDR_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return FDR_OK;
}
__attribute__((visibility("default"))) int fdr_host_free(void* ptr) {
    if (ptr) FDR_CUDA(cudaFreeHost(ptr));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_create(fdr_plan** plan, int rows, int cols, int channels, int max_images, int device) {
    if (!plan) return set_error(FDR_E_INVALID, "plan is NULL");
    *plan = nullptr;
    if (rows < 1 || cols < 1 || channels < 1 || max_images < 1)
        return set_error(FDR_E_INVALID, "rows=%d cols=%d channels=%d max_images=%d must all be >= 1", rows, cols, channels, max_images);
    if (rows > 16384 || cols > 16384) return set_error(FDR_E_INVALID, "padded size above 16384 is not supported (%dx%d)", rows, cols);
    FDR_TRY(ensure_device(device));
    fdr_plan* p = new (std::nothrow) fdr_plan();
    if (!p) return set_error(FDR_E_NOMEM, "out of host memory");
    p->device = device;
    p->H = rows;
    p->W = cols;
    p->C = channels;
    p->max_images = max_images;
    p->Rp = next_pow2(rows);
    p->Cp = next_pow2(cols);
    {
        {
            ColPassArgs probe{};
            probe.n = p->Rp;
            probe.pitch = p->Cp;
            probe.mode = COL_WIENER;
            const char* cs = getenv("FDR_COL_SPLIT");
            p->col_split = col_split_applicable(probe) && !(cs && atoi(cs) == 0);
        }
        const char* ln = getenv("FDR_LANES");
        if (ln && atoi(ln) >= 1 && atoi(ln) <= 4) p->lanes = atoi(ln);
        {
            // half-plane path for the odd colour plane: from FDR_HALF_MIN_PIXELS padded pixels up (below that an image
            // is launch-bound and the extra launches cost more than the saved traffic); FDR_HALF=0 disables it
            const char* hv = getenv("FDR_HALF");
            const char* hm = getenv("FDR_HALF_MIN_PIXELS");
            const long long min_px = (hm && atoll(hm) >= 0) ? atoll(hm) : (1LL << 22);
            p->half_ok = !(hv && atoi(hv) == 0) && p->Cp >= FDR_HALF_MIN_N && p->Rp >= 2 && (long long)p->Rp * p->Cp >= min_px;
            ColPassArgs probe{};
            probe.n = p->Rp;
            probe.pitch = p->Cp / 2;
            probe.mode = COL_WIENER;
            const char* cs = getenv("FDR_COL_SPLIT");
            p->col_split_half = p->half_ok && col_split_applicable(probe) && !(cs && atoi(cs) == 0);
        }
    }
    cudaError_t e = get_twiddles(p->Cp, &p->tw_rows);
    if (e == cudaSuccess) e = get_twiddles(p->Rp, &p->tw_cols);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&p->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_last, cudaEventDisableTiming);
    if (e == cudaSuccess && p->half_ok) e = cudaStreamCreateWithFlags(&p->s_half, cudaStreamNonBlocking);
    if (e == cudaSuccess && p->half_ok) e = cudaEventCreateWithFlags(&p->ev_half_fork, cudaEventDisableTiming);
    if (e == cudaSuccess && p->half_ok) e = cudaEventCreateWithFlags(&p->ev_half_join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        const int rc = set_error(FDR_E_CUDA, "plan twiddle/stream/event creation: %s", cudaGetErrorString(e));
        fdr_plan_destroy(p);  // releases whatever was created
        return rc;
    }
    *plan = p;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_destroy(fdr_plan* p) {
    if (!p) return FDR_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    p->spec.release();
    p->raw.release();
    p->mm.release();
    p->ss.release();
    p->mmf.release();
    p->wb_sums.release();
    p->wiener.release();
    p->wiener_nat.release();
    p->wiener_tiled.release();
    p->wiener_half.release();
    p->wiener_half_tiled.release();
    p->wiener_nyq.release();
    p->psf.release();
    p->d_in_u8.release();
    p->d_out_u8.release();
    p->d_in_f32.release();
    p->d_out_f32.release();
    p->h_u8_in.release();
    p->h_u8_out.release();
    p->h_f32_in.release();
    p->h_f32_out.release();
    for (int i = 0; i < 8; ++i)
        if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    for (auto& r : p->krecs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : p->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_cmp[i]) cudaEventDestroy(p->ev_cmp[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    }
    for (int i = 0; i < 4; ++i) {
        if (p->lane_stream[i]) cudaStreamDestroy(p->lane_stream[i]);
        if (p->lane_join[i]) cudaEventDestroy(p->lane_join[i]);
    }
    if (p->lane_fork) cudaEventDestroy(p->lane_fork);
    if (p->ev_last) cudaEventDestroy(p->ev_last);
    if (p->ev_half_fork) cudaEventDestroy(p->ev_half_fork);
    if (p->ev_half_join) cudaEventDestroy(p->ev_half_join);
    if (p->s_half) cudaStreamDestroy(p->s_half);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_padded_size(const fdr_plan* p, int* pr, int* pc) {
    if (!p) return set_error(FDR_E_INVALID, "plan is NULL");
    if (pr) *pr = p->Rp;
    if (pc) *pc = p->Cp;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_set_white_balance(fdr_plan* p, int enabled) {
    if (!p) return set_error(FDR_E_INVALID, "plan is NULL");
    if (enabled && p->C != 3) return set_error(FDR_E_INVALID, "white balance needs 3-channel (BGR) images");
    p->white_balance = enabled != 0;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_half_plane(const fdr_plan* p, int* enabled) {
    if (!p || !enabled) return set_error(FDR_E_INVALID, "bad arguments");
    *enabled = p->half_ok ? 1 : 0;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_set_chunk_images(fdr_plan* p, int images) {
    if (!p || images < 0) return set_error(FDR_E_INVALID, "bad plan or chunk size");
    p->chunk_images_user = images;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_set_psf_host(fdr_plan* p, const float* psf, int psf_rows, int psf_cols, size_t stride_bytes, float K) {
    if (!p || !psf || psf_rows < 1 || psf_cols < 1) return set_error(FDR_E_INVALID, "bad PSF arguments");
    if (stride_bytes == 0) stride_bytes = (size_t)psf_cols * sizeof(float);
    FDR_TRY(ensure_device(p->device));
    FDR_TRY(p->psf.ensure((size_t)psf_rows * psf_cols));
    FDR_CUDA(cudaMemcpy2DAsync(p->psf.p, (size_t)psf_cols * sizeof(float), psf, stride_bytes, (size_t)psf_cols * sizeof(float),
                               psf_rows, cudaMemcpyHostToDevice, p->stream));
    FDR_CUDA(cudaStreamSynchronize(p->stream));
    p->psf_rows = psf_rows;
    p->psf_cols = psf_cols;
    p->K = K;
    p->have_wiener = false;
    return build_wiener(p);
}

__attribute__((visibility("default"))) int fdr_plan_set_psf_motion(fdr_plan* p, int length, double angle_deg, float K) {
    if (!p || length < 1) return set_error(FDR_E_INVALID, "bad PSF length %d", length);
    FDR_TRY(ensure_device(p->device));
    FDR_TRY(p->psf.ensure((size_t)length * length));
    FDR_CUDA(launch_motion_psf(p->psf.p, length, motion_affine(length, angle_deg), p->stream));
    p->psf_rows = p->psf_cols = length;
    p->K = K;
    p->have_wiener = false;
    return build_wiener(p);
}

__attribute__((visibility("default"))) int fdr_plan_get_psf_host(const fdr_plan* p, float* out, int capacity, int* rows, int* cols) {
    if (!p) return set_error(FDR_E_INVALID, "plan is NULL");
    if (p->psf_rows == 0) return set_error(FDR_E_STATE, "no PSF set");
    if (rows) *rows = p->psf_rows;
    if (cols) *cols = p->psf_cols;
    if (out) {
        if (capacity < p->psf_rows * p->psf_cols) return set_error(FDR_E_INVALID, "PSF buffer too small");
        FDR_CUDA(cudaSetDevice(p->device));
        FDR_CUDA(cudaMemcpy(out, p->psf.p, sizeof(float) * (size_t)p->psf_rows * p->psf_cols, cudaMemcpyDeviceToHost));
    }
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_get_wiener_host(const fdr_plan* p, float* wf) {
    if (!p || !wf) return set_error(FDR_E_INVALID, "bad arguments");
    if (!p->have_wiener) return set_error(FDR_E_STATE, "no PSF set");
    FDR_CUDA(cudaSetDevice(p->device));
    if (p->col_split) {
        // stored with digit-swapped rows: row M*k1 + k2 holds frequency k1 + (Rp/M)*k2 (layout conversion only)
        std::vector<float2> tmp(p->plane_elems());
        FDR_CUDA(cudaMemcpy(tmp.data(), p->wiener.p, sizeof(float2) * p->plane_elems(), cudaMemcpyDeviceToHost));
        float2* out = reinterpret_cast<float2*>(wf);
        ColPassArgs probe{};
        probe.n = p->Rp;
        probe.pitch = p->Cp;
        probe.mode = COL_WIENER;
        const int M = col_split_block_len(probe), n1 = p->Rp / M;
        for (int k1 = 0; k1 < n1; ++k1)
            for (int k2 = 0; k2 < M; ++k2)
                memcpy(out + (size_t)(k1 + n1 * k2) * p->Cp, tmp.data() + (size_t)(M * k1 + k2) * p->Cp, sizeof(float2) * p->Cp);
        return FDR_OK;
    }
    FDR_CUDA(cudaMemcpy(wf, p->wiener.p, sizeof(float2) * p->plane_elems(), cudaMemcpyDeviceToHost));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_restore_images_device_u8(fdr_plan* p, const void* d_in, void* d_out, int n_images, void* stream) {
    if (!p || !d_in || !d_out || n_images < 0) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_TRY(ensure_device(p->device));
    InputDesc in;
    in.mode = ROW_IN_PAIR_U8;
    in.u8 = static_cast<const uint8_t*>(d_in);
    return restore_units_device(p, in, nullptr, static_cast<uint8_t*>(d_out), (long long)n_images * p->C,
                                stream ? static_cast<cudaStream_t>(stream) : p->stream);
}

__attribute__((visibility("default"))) int fdr_restore_planes_device_f32(fdr_plan* p, const void* d_in, void* d_out_f32, void* d_out_u8, int n_planes, void* stream) {
    if (!p || !d_in || n_planes < 0 || (!d_out_f32 && !d_out_u8)) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_TRY(ensure_device(p->device));
    InputDesc in;
    in.mode = ROW_IN_PAIR_F32;
    in.f32 = static_cast<const float*>(d_in);
    in.unit_stride = (long long)p->H * p->W;
    in.row_stride = p->W;
    return restore_units_device(p, in, static_cast<float*>(d_out_f32), static_cast<uint8_t*>(d_out_u8), n_planes,
                                stream ? static_cast<cudaStream_t>(stream) : p->stream);
}

__attribute__((visibility("default"))) int fdr_restore_planes_host_f32(fdr_plan* p, const float* const* in_planes, size_t in_stride, float* const* out_planes,
                                size_t out_stride, int n_planes) {
    if (!p || !in_planes || !out_planes || n_planes < 1) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_TRY(ensure_device(p->device));
    const size_t HW = (size_t)p->H * p->W, rowb = (size_t)p->W * sizeof(float);
    if (in_stride == 0) in_stride = rowb;
    if (out_stride == 0) out_stride = rowb;
    if (in_stride < rowb || out_stride < rowb) return set_error(FDR_E_INVALID, "row stride smaller than a row");
    for (int u = 0; u < n_planes; ++u)
        if (!in_planes[u] || !out_planes[u]) return set_error(FDR_E_INVALID, "plane %d is NULL", u);
    cudaStream_t s = p->stream;
    for (int i = 0; i < 6; ++i) p->profile_ms[i] = 0.f;
    {
        ScopedTimer t(p, 0, s, &p->profile_ms[0]);
        FDR_TRY(p->d_in_f32.ensure(HW * n_planes));
        FDR_TRY(p->d_out_f32.ensure(HW * n_planes));
        t.stop();
    }
    {
        // straight from the caller's planes (strided 2-D copies; no host staging pass)
        ScopedTimer t(p, 1, s, &p->profile_ms[1]);
        for (int u = 0; u < n_planes; ++u)
            FDR_CUDA(cudaMemcpy2DAsync(p->d_in_f32.p + u * HW, rowb, in_planes[u], in_stride, rowb, p->H, cudaMemcpyHostToDevice, s));
        t.stop();
    }
    {
        ScopedTimer t(p, 2, s, &p->profile_ms[3]);
        InputDesc in;
        in.mode = ROW_IN_PAIR_F32;
        in.f32 = p->d_in_f32.p;
        in.unit_stride = (long long)HW;
        in.row_stride = p->W;
        FDR_TRY(restore_units_device(p, in, p->d_out_f32.p, nullptr, n_planes, s));
        t.stop();
    }
    {
        ScopedTimer t(p, 3, s, &p->profile_ms[4]);
        for (int u = 0; u < n_planes; ++u)
            FDR_CUDA(cudaMemcpy2DAsync(out_planes[u], out_stride, p->d_out_f32.p + u * HW, rowb, rowb, p->H, cudaMemcpyDeviceToHost, s));
        t.stop();
    }
    return FDR_OK;
}

// Whole images through host buffers, pipelined: the batch is cut into chunks and the H2D copy of
// chunk k+1, the restoration of chunk k and the D2H copy of chunk k-1 run concurrently on three
// streams (PCIe is full duplex), with double-buffered device staging.  The reference does
// memcpy -> H2D -> compute -> D2H -> sync serially per channel (fft_gpu.cu:347-349,373-374).
__attribute__((visibility("default"))) int fdr_restore_images_host_u8(fdr_plan* p, const uint8_t* in_images, uint8_t* out_images, int n_images) {
    if (!p || !in_images || !out_images || n_images < 1) return set_error(FDR_E_INVALID, "bad arguments");
    if (!p->have_wiener) return set_error(FDR_E_STATE, "no PSF set: call fdr_plan_set_psf_* first");
    FDR_TRY(ensure_device(p->device));
    const size_t img_bytes = (size_t)p->H * p->W * p->C;
    const int chunk = p->host_chunk_images(n_images);
    const size_t chunk_bytes = img_bytes * chunk;
    for (int i = 0; i < 6; ++i) p->profile_ms[i] = 0.f;
    if (!p->s_in) {
        FDR_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        FDR_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            FDR_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            FDR_CUDA(cudaEventCreateWithFlags(&p->ev_cmp[i], cudaEventDisableTiming));
            FDR_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        }
    }
    {
        ScopedTimer t(p, 0, p->stream, &p->profile_ms[0]);
        FDR_TRY(p->d_in_u8.ensure(2 * chunk_bytes));
        FDR_TRY(p->d_out_u8.ensure(2 * chunk_bytes));
        FDR_TRY(ensure_workspace(p, chunk * p->C));
        t.stop();
    }
    cudaStream_t s_in = p->s_in, s_cmp = p->stream, s_out = p->s_out;
    FDR_CUDA(cudaEventRecord(p->ev[2], s_cmp));  // wall span of the pipelined region (bucket 4: compute)
    long long launches = 0;
    int k = 0;
    for (int first = 0; first < n_images; first += chunk, ++k) {
        const int n = (n_images - first < chunk) ? (n_images - first) : chunk;
        const int b = k & 1;
        const size_t off = img_bytes * first, bytes = img_bytes * n;
        uint8_t* din = p->d_in_u8.p + (size_t)b * chunk_bytes;
        uint8_t* dout = p->d_out_u8.p + (size_t)b * chunk_bytes;
        if (k >= 2) FDR_CUDA(cudaStreamWaitEvent(s_in, p->ev_cmp[b], 0));  // chunk k-2 no longer reads din
        FDR_CUDA(cudaMemcpyAsync(din, in_images + off, bytes, cudaMemcpyHostToDevice, s_in));
        FDR_CUDA(cudaEventRecord(p->ev_in[b], s_in));
        FDR_CUDA(cudaStreamWaitEvent(s_cmp, p->ev_in[b], 0));
        if (k >= 2) FDR_CUDA(cudaStreamWaitEvent(s_cmp, p->ev_out[b], 0));  // chunk k-2's D2H has drained dout
        InputDesc in;
        in.mode = ROW_IN_PAIR_U8;
        in.u8 = din;
        FDR_TRY(restore_units_device(p, in, nullptr, dout, (long long)n * p->C, s_cmp));
        launches += p->launches;
        FDR_CUDA(cudaEventRecord(p->ev_cmp[b], s_cmp));
        FDR_CUDA(cudaStreamWaitEvent(s_out, p->ev_cmp[b], 0));
        FDR_CUDA(cudaMemcpyAsync(out_images + off, dout, bytes, cudaMemcpyDeviceToHost, s_out));
        FDR_CUDA(cudaEventRecord(p->ev_out[b], s_out));
    }
    FDR_CUDA(cudaStreamSynchronize(s_out));
    FDR_CUDA(cudaStreamSynchronize(s_in));
    FDR_CUDA(cudaEventRecord(p->ev[3], s_cmp));
    FDR_CUDA(cudaEventSynchronize(p->ev[3]));
    float ms = 0.f;
    FDR_CUDA(cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]));
    p->profile_ms[3] = ms;  // H2D, compute and D2H overlap: reported as one bucket
    p->launches = launches;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_last_minmax_host(fdr_plan* p, float* minmax, int capacity_planes) {
    if (!p || !minmax) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_CUDA(cudaSetDevice(p->device));
    FDR_CUDA(cudaDeviceSynchronize());
    int n = p->last_units < capacity_planes ? p->last_units : capacity_planes;
    if (n > 0)
        FDR_CUDA(cudaMemcpy(minmax, p->mmf.p + (size_t)p->last_lane * p->ws_units * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_get_profile(const fdr_plan* p, float ms[6]) {
    if (!p || !ms) return set_error(FDR_E_INVALID, "bad arguments");
    for (int i = 0; i < 6; ++i) ms[i] = p->profile_ms[i];
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_last_launch_count(const fdr_plan* p, long long* launches) {
    if (!p || !launches) return set_error(FDR_E_INVALID, "bad arguments");
    *launches = p->launches;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_set_kernel_timing(fdr_plan* p, int enabled) {
    if (!p) return set_error(FDR_E_INVALID, "plan is NULL");
    p->ktiming = enabled != 0;
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_get_kernel_timing(fdr_plan* p, double total_ms[4], long long launches[4], double bytes[4]) {
    if (!p || !total_ms || !launches || !bytes) return set_error(FDR_E_INVALID, "bad arguments");
    for (int i = 0; i < 4; ++i) {
        total_ms[i] = 0;
        launches[i] = 0;
        bytes[i] = 0;
    }
    FDR_CUDA(cudaSetDevice(p->device));
    for (auto& r : p->krecs) {
        FDR_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        FDR_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        total_ms[r.kind] += ms;
        launches[r.kind] += 1;
        bytes[r.kind] += r.bytes;
        p->ev_pool.push_back(r.a);
        p->ev_pool.push_back(r.b);
    }
    p->krecs.clear();
    return FDR_OK;
}

// Timing probe: runs one pass `reps` times on a workspace of `npairs` plane pairs and returns the
// mean device time.  pass: 1 = rows forward (u8 in), 2 = columns, 3 = rows inverse + min/max.
// variant (pass 2): 0 = Wiener, default dispatch; 1 = Wiener, plain-load kernel; 2 = one forward FFT;
// 3 = load + store only; 4 = Wiener, TMA kernel with one tile per CTA; 5 = Wiener, persistent pipelined TMA kernel;
// 6 = Wiener, TMA kernel on the 64-points-per-thread core; 7 / 8 = its transfer-only probes (2 / 3 tile transfers, no FFT);
// 9 = the wide core in the persistent pipelined form; 10 = wide core at 4096 with 4-column tiles (one CTA per SM).
__attribute__((visibility("default"))) int fdr_plan_time_pass(fdr_plan* p, int pass, int variant, int npairs, int reps, float* ms_avg) {
    if (!p || !ms_avg || npairs < 1 || reps < 1) return set_error(FDR_E_INVALID, "bad arguments");
    if (!p->have_wiener) return set_error(FDR_E_STATE, "no PSF set");
    FDR_TRY(ensure_device(p->device));
    const int nu = 2 * npairs;
    FDR_TRY(ensure_workspace(p, nu));
    const int n_img = (nu + 2) / 3;  // interleaved BGR images, as the real pipeline sees them
    FDR_TRY(p->d_in_u8.ensure((size_t)n_img * 3 * p->H * p->W));
    cudaStream_t s = p->stream;
    FDR_CUDA(launch_synth_u8(p->d_in_u8.p, 12345u, 0, n_img, 3, (long long)p->H * p->W, 0, (long long)p->H * p->W, s));
    RowPassArgs r1{};
    r1.n = p->Cp; r1.nrows = p->H; r1.npairs = npairs; r1.in_mode = ROW_IN_PAIR_U8; r1.out_mode = ROW_OUT_COMPLEX;
    r1.in_u8 = p->d_in_u8.p; r1.channels = 3; r1.img_rows = p->H; r1.img_cols = p->W; r1.units_total = nu;
    r1.cout = p->spec.p; r1.cplane = (long long)p->plane_elems(); r1.tw = p->tw_rows;
    ColPassArgs c2{};
    c2.n = p->Rp; c2.pitch = p->Cp; c2.npairs = npairs; c2.rows_valid = p->H; c2.data = p->spec.p;
    c2.cplane = (long long)p->plane_elems(); c2.wiener = p->wiener.p; c2.K = p->K; c2.tw = p->tw_cols;
    c2.wiener_tiled = getenv("FDR_NO_WIENER_TILED") ? nullptr : p->wiener_tiled.p;
    c2.mode = variant == 2 ? COL_FFT : variant == 3 ? COL_COPY : COL_WIENER;
    c2.col_variant = (variant == 1) ? 1 : (variant == 4) ? 2 : (variant == 5) ? 3 : (variant == 6) ? 4 : (variant == 7) ? 5 : (variant == 8) ? 6 : (variant == 9) ? 7 : (variant == 10) ? 8 : 0;
    RowPassArgs r3{};
    r3.n = p->Cp; r3.nrows = p->Rp; r3.npairs = npairs; r3.in_mode = ROW_IN_COMPLEX; r3.out_mode = ROW_OUT_REAL_PAIR;
    r3.cin = p->spec.p; r3.cplane = (long long)p->plane_elems(); r3.units_total = nu; r3.raw = p->raw.p;
    r3.raw_unit_stride = (long long)p->H * p->W; r3.raw_rows = p->H; r3.raw_cols = p->W; r3.minmax = p->mm.p; r3.local_units = nu;
    r3.tw = p->tw_rows;
    FDR_CUDA(launch_row_pass(r1, s));  // realistic data in the workspace
    FDR_CUDA(launch_minmax_reset(p->mm.p, nu, s));
    auto run = [&]() -> cudaError_t {
        if (pass == 1) return launch_row_pass(r1, s);
        if (pass == 2 && variant >= 100) return launch_tma_copy_probe(c2, variant - 100, s);  // 100 + box columns
        if (pass == 2) return (p->col_split && c2.mode == COL_WIENER) ? launch_col_split(c2, s, nullptr) : launch_col_pass(c2, s);
        return launch_row_pass(r3, s);
    };
    FDR_CUDA(run());
    FDR_CUDA(cudaEventRecord(p->ev[0], s));
    for (int i = 0; i < reps; ++i) FDR_CUDA(run());
    FDR_CUDA(cudaEventRecord(p->ev[1], s));
    FDR_CUDA(cudaEventSynchronize(p->ev[1]));
    float ms = 0.f;
    FDR_CUDA(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]));
    *ms_avg = ms / reps;
    return FDR_OK;
}

static int spectrum_host(fdr_plan* p, const float* plane, size_t stride, float* out, int col_mode) {
    if (!p || !plane || !out) return set_error(FDR_E_INVALID, "bad arguments");
    if (col_mode == COL_FILTER && !p->have_wiener) return set_error(FDR_E_STATE, "no PSF set");
    FDR_TRY(ensure_device(p->device));
    const size_t HW = (size_t)p->H * p->W, rowb = (size_t)p->W * sizeof(float);
    if (stride == 0) stride = rowb;
    cudaStream_t s = p->stream;
    FDR_TRY(p->d_in_f32.ensure(HW));
    FDR_TRY(p->spec.ensure(p->plane_elems()));
    FDR_CUDA(cudaMemcpy2DAsync(p->d_in_f32.p, rowb, plane, stride, rowb, p->H, cudaMemcpyHostToDevice, s));
    RowPassArgs r{};
    r.n = p->Cp;
    r.nrows = p->H;
    r.npairs = 1;
    r.in_mode = ROW_IN_PAIR_F32;
    r.out_mode = ROW_OUT_COMPLEX;
    r.in_f32 = p->d_in_f32.p;
    r.in_unit_stride = (long long)HW;
    r.in_row_stride = p->W;
    r.channels = 1;
    r.img_rows = p->H;
    r.img_cols = p->W;
    r.units_total = 1;
    r.cout = p->spec.p;
    r.cplane = (long long)p->plane_elems();
    r.tw = p->tw_rows;
    FDR_CUDA(launch_row_pass(r, s));
    ColPassArgs c{};
    c.n = p->Rp;
    c.pitch = p->Cp;
    c.npairs = 1;
    c.mode = col_mode;
    c.rows_valid = p->H;
    c.data = p->spec.p;
    c.cplane = (long long)p->plane_elems();
    c.wiener = p->wiener.p;
    c.tw = p->tw_cols;
    if (col_mode == COL_FILTER && p->col_split) {
        if (!p->wiener_nat.p) {
            // the row pass above wrote the image spectrum into spec; build the natural-order factor first
            FDR_TRY(build_wiener_into(p, p->wiener_nat, false));
            FDR_CUDA(launch_row_pass(r, s));
        }
        c.wiener = p->wiener_nat.p;
    }
    FDR_CUDA(launch_col_pass(c, s));
    FDR_CUDA(cudaMemcpyAsync(out, p->spec.p, sizeof(float2) * p->plane_elems(), cudaMemcpyDeviceToHost, s));
    FDR_CUDA(cudaStreamSynchronize(s));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_plan_forward_spectrum_host(fdr_plan* p, const float* plane, size_t stride, float* G) {
    return spectrum_host(p, plane, stride, G, COL_FFT);
}
__attribute__((visibility("default"))) int fdr_plan_filtered_spectrum_host(fdr_plan* p, const float* plane, size_t stride, float* F) {
    return spectrum_host(p, plane, stride, F, COL_FILTER);
}

// ---- building blocks ----------------------------------------------------------------------

// rows x n complex, transform every row (in place on the device buffer d).
static int rows_device(float2* d, float2* tmp, int rows, int n, int inverse, cudaStream_t s) {
    if (is_pow2(n)) {
        RowPassArgs r{};
        r.n = n;
        r.nrows = rows;
        r.npairs = 1;
        r.in_mode = ROW_IN_COMPLEX;
        r.out_mode = ROW_OUT_COMPLEX;
        r.conj = inverse ? 1 : 0;
        FDR_CUDA(get_twiddles(n, &r.tw));
        r.cin = d;
        r.cout = d;
        r.cplane = (long long)rows * n;
        FDR_CUDA(launch_row_pass(r, s));
    } else {
        FDR_CUDA(launch_dft_naive(d, tmp, n, 1, rows, n, inverse, s));
        FDR_CUDA(cudaMemcpyAsync(d, tmp, sizeof(float2) * (size_t)rows * n, cudaMemcpyDeviceToDevice, s));
    }
    return FDR_OK;
}

static int cols_device(float2* d, float2* tmp, int rows, int cols, int inverse, cudaStream_t s) {
    if (is_pow2(rows)) {
        ColPassArgs c{};
        c.n = rows;
        c.pitch = cols;
        c.npairs = 1;
        c.mode = COL_FFT;
        c.conj = inverse ? 1 : 0;
        FDR_CUDA(get_twiddles(rows, &c.tw));
        c.rows_valid = rows;
        c.data = d;
        c.cplane = (long long)rows * cols;
        FDR_CUDA(launch_col_pass(c, s));
    } else {
        FDR_CUDA(launch_dft_naive(d, tmp, rows, cols, cols, 1, inverse, s));
        FDR_CUDA(cudaMemcpyAsync(d, tmp, sizeof(float2) * (size_t)rows * cols, cudaMemcpyDeviceToDevice, s));
    }
    return FDR_OK;
}

static int transform_host(float* data, int rows, int cols, int inverse, bool do_rows, bool do_cols) {
    if (!data || rows < 1 || cols < 1) return set_error(FDR_E_INVALID, "bad arguments");
    if ((is_pow2(cols) && cols > 16384) || (is_pow2(rows) && do_cols && rows > 16384))
        return set_error(FDR_E_INVALID, "power-of-two lengths above 16384 are not supported");
    int dev = 0;
    FDR_CUDA(cudaGetDevice(&dev));
    FDR_TRY(ensure_device(dev));
    const size_t n = (size_t)rows * cols;
    float2 *d = nullptr, *tmp = nullptr;
    FDR_CUDA(cudaMalloc(&d, n * sizeof(float2)));
    const bool need_tmp = (do_rows && !is_pow2(cols)) || (do_cols && !is_pow2(rows));
    int rc = FDR_OK;
    cudaError_t e = cudaSuccess;
    if (need_tmp) e = cudaMalloc(&tmp, n * sizeof(float2));
    if (e != cudaSuccess) rc = set_error(FDR_E_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    if (rc == FDR_OK) {
        e = cudaMemcpy(d, data, n * sizeof(float2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "H2D: %s", cudaGetErrorString(e));
    }
    if (rc == FDR_OK && do_rows) rc = rows_device(d, tmp, rows, cols, inverse, 0);
    if (rc == FDR_OK && do_cols) rc = cols_device(d, tmp, rows, cols, inverse, 0);
    if (rc == FDR_OK) {
        e = cudaMemcpy(data, d, n * sizeof(float2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "D2H: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    if (tmp) cudaFree(tmp);
    return rc;
}

__attribute__((visibility("default"))) int fdr_dft2d_host(float* data, int rows, int cols, int inverse) { return transform_host(data, rows, cols, inverse, true, true); }

__attribute__((visibility("default"))) int fdr_fft_radix2_host(float* data, int n, int inverse) {
    if (!is_pow2(n)) return set_error(FDR_E_INVALID, "fft_radix2 needs a power-of-two length, got %d", n);
    return transform_host(data, 1, n, inverse, true, false);
}

__attribute__((visibility("default"))) int fdr_dft_naive_host(float* data, int n, int inverse) {
    if (!data || n < 1) return set_error(FDR_E_INVALID, "bad arguments");
    int dev = 0;
    FDR_CUDA(cudaGetDevice(&dev));
    FDR_TRY(ensure_device(dev));
    float2 *d = nullptr, *o = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(float2) * n);
    if (e == cudaSuccess) e = cudaMalloc(&o, sizeof(float2) * n);
    if (e == cudaSuccess) e = cudaMemcpy(d, data, sizeof(float2) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_dft_naive(d, o, n, 1, 1, n, inverse, 0);
    if (e == cudaSuccess) e = cudaMemcpy(data, o, sizeof(float2) * n, cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (o) cudaFree(o);
    if (e != cudaSuccess) return set_error(FDR_E_CUDA, "naive DFT: %s", cudaGetErrorString(e));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_transform_rows_host(float* data, int rows, int n, int inverse) { return transform_host(data, rows, n, inverse, true, false); }

// The drivers' post stage alone (gpu.cpp:123-134): restored planes + original planes (B, G, R; f32 in
// [0,1]; rows x cols, contiguous) -> white-balanced 8-bit BGR image.  Runs on the current device.
__attribute__((visibility("default"))) int fdr_white_balance_pack_host(const float* const* restored_planes, const float* const* original_planes, int rows,
                                                                 int cols, uint8_t* out_bgr) {
    if (!restored_planes || !original_planes || !out_bgr || rows < 1 || cols < 1) return set_error(FDR_E_INVALID, "bad arguments");
    const size_t HW = (size_t)rows * cols;
    float *d_r = nullptr, *d_o = nullptr;
    double* d_s = nullptr;
    float2* d_ss = nullptr;
    uint8_t* d_out = nullptr;
    int rc = FDR_OK;
    cudaError_t e = cudaMalloc(&d_r, 3 * HW * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_o, 3 * HW * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_s, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_ss, 3 * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, 3 * HW);
    const float2 ident[3] = {{1.f, 0.f}, {1.f, 0.f}, {1.f, 0.f}};
    if (e == cudaSuccess) e = cudaMemcpy(d_ss, ident, sizeof(ident), cudaMemcpyHostToDevice);
    for (int c = 0; c < 3 && e == cudaSuccess; ++c) {
        if (!restored_planes[c] || !original_planes[c]) {
            rc = set_error(FDR_E_INVALID, "plane %d is NULL", c);
            break;
        }
        e = cudaMemcpy(d_r + c * HW, restored_planes[c], HW * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_o + c * HW, original_planes[c], HW * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (rc == FDR_OK && e == cudaSuccess) e = launch_white_balance_pack_u8(d_r, (long long)HW, d_ss, nullptr, d_o, (long long)HW, d_s, d_out, 1, rows, cols, 0);
    if (rc == FDR_OK && e == cudaSuccess) e = cudaMemcpy(out_bgr, d_out, 3 * HW, cudaMemcpyDeviceToHost);
    cudaFree(d_r);
    cudaFree(d_o);
    cudaFree(d_s);
    cudaFree(d_ss);
    cudaFree(d_out);
    if (rc != FDR_OK) return rc;
    if (e != cudaSuccess) return set_error(FDR_E_CUDA, "white balance: %s", cudaGetErrorString(e));
    return FDR_OK;
}

// Plain copies for harnesses that hold raw device pointers (kind: 0 = H2D, 1 = D2H, 2 = D2D).
__attribute__((visibility("default"))) int fdr_memcpy(void* dst, const void* src, size_t bytes, int kind) {
    if (!dst || !src || kind < 0 || kind > 2) return set_error(FDR_E_INVALID, "bad arguments");
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    FDR_CUDA(cudaMemcpy(dst, src, bytes, k));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_motion_psf_host(int length, double angle_deg, float* psf_out) {
    if (length < 1 || !psf_out) return set_error(FDR_E_INVALID, "bad PSF arguments");
    float* d = nullptr;
    FDR_CUDA(cudaMalloc(&d, sizeof(float) * (size_t)length * length));
    cudaStream_t st = nullptr;   // own non-blocking stream: nothing of this library runs on the legacy default stream
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = launch_motion_psf(d, length, motion_affine(length, angle_deg), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(psf_out, d, sizeof(float) * (size_t)length * length, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (st) cudaStreamDestroy(st);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(FDR_E_CUDA, "motion PSF: %s", cudaGetErrorString(e));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_synth_images_device_u8(void* d_out, uint32_t seed, long long first_image, int n_images, int channels, int rows,
                               int cols, void* stream) {
    if (!d_out || n_images < 0 || channels < 1 || rows < 1 || cols < 1) return set_error(FDR_E_INVALID, "bad arguments");
    if (n_images == 0) return FDR_OK;
    FDR_CUDA(launch_synth_u8(static_cast<uint8_t*>(d_out), seed, first_image, n_images, channels, (long long)rows * cols, 0,
                             (long long)rows * cols, static_cast<cudaStream_t>(stream)));
    return FDR_OK;
}

__attribute__((visibility("default"))) int fdr_l2_flush_device(void* d_scratch, size_t bytes, void* stream) {
    if (!d_scratch || bytes < 16) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_CUDA(launch_l2_flush(d_scratch, bytes, static_cast<cudaStream_t>(stream)));
    return FDR_OK;
}

}  // extern "C"
