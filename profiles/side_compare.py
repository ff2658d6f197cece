"""Side comparisons on the same B200 (SURVEY.md 8d "CPU baseline beside it"), NOT product code:
  * the reference's own gpu mode (fft/fft_gpu.cu compiled unmodified for sm_100a, oracle/_ref/libref_gpu.so),
    timed as gpu.cpp does (wall clock around the second wienerDeblur_RGB_optimized call);
  * a cuFFT pipeline through torch.fft (fft2 -> multiply -> ifft2 -> real -> amin/amax -> normalise -> u8), device resident;
  * this repo through the same boundaries.
Usage: python profiles/side_compare.py  > profiles/r1/side_compare.txt"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_fdr, load_oracle  # noqa: E402

fdr = load_fdr()
O = load_oracle()
K = 0.01
refgpu = None
p = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
if os.path.exists(p):
    refgpu = C.CDLL(p)
    refgpu.ref_gpu_restore.restype = C.c_double
    refgpu.ref_gpu_restore.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float]


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


print("== host-buffer boundary: 3 f32 planes in, 3 normalised f32 planes out (fft.hpp:33), wall-clock ms ==")
for name, (H, W, S, ang) in {"car 330x640": (330, 640, 40, 45.0), "cat 782x1920": (782, 1920, 50, 30.0),
                             "synthetic 2048x2048": (2048, 2048, 50, 30.0), "synthetic 4096x4096": (4096, 4096, 50, 30.0)}.items():
    planes = np.stack([O.synth_image_u8(3, 0, H, W)[c].astype(np.float32) / np.float32(255) for c in range(3)])
    psf = O.port().motion_psf(S, ang)
    line = "%-22s" % name
    if refgpu is not None:
        buf = planes.copy()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)  # the reference prints its profile block on every call
        try:
            refgpu.ref_gpu_restore(0, fp(buf), 3, H, W, fp(psf), S, S, K)  # warm-up (gpu.cpp:96)
            ts = []
            for _ in range(3):
                buf = planes.copy()
                ts.append(refgpu.ref_gpu_restore(0, fp(buf), 3, H, W, fp(psf), S, S, K))
            buf = planes.copy()
            tn = refgpu.ref_gpu_restore(1, fp(buf), 3, H, W, fp(psf), S, S, K)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
        line += " | reference gpu[optimize] %8.2f ms, gpu(naive) %8.2f ms" % (min(ts), tn)
    with fdr.Plan(H, W, 1) as plan:
        plan.set_psf(psf, K)
        plan.restore_planes(list(planes))
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            plan.restore_planes(list(planes))
            ts.append((time.perf_counter() - t0) * 1e3)
        prof = plan.profile()
    line += " | this repo %8.2f ms (GPU compute bucket %.3f ms)" % (min(ts), prof[3])
    print(line)

try:
    import torch
    print("== device-resident, 64 x 2048x2048x3 u8 -> u8: cuFFT pipeline via torch.fft vs this repo ==")
    H = W = 2048
    B = 64
    dev = torch.device("cuda", 0)
    d_in = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(d_in)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    fdr.synth_images_device_u8(d_in.data_ptr(), 0xF17E0003, 0, B, 3, H, W, st.cuda_stream)
    with fdr.Plan(H, W, 3, B) as plan:
        plan.set_psf_motion(50, 30.0, K)
        wf = torch.from_numpy(plan.get_wiener()).to(dev)
        for _ in range(3):
            plan.restore_images_device_u8(d_in.data_ptr(), d_out.data_ptr(), B, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            plan.restore_images_device_u8(d_in.data_ptr(), d_out.data_ptr(), B, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ours = e0.elapsed_time(e1) / 5

    def cufft_pipeline(chunk=8):
        for b0 in range(0, B, chunk):
            x = d_in[b0:b0 + chunk].permute(0, 3, 1, 2).to(torch.float32) * (1.0 / 255.0)  # (n,3,H,W)
            z = torch.complex(x[:, 0::2][:, :1], x[:, 1:2])                                 # pack B + iG
            zs = torch.cat([z, torch.complex(x[:, 2:3], torch.zeros_like(x[:, 2:3]))], 1)   # and R + i0
            f = torch.fft.ifft2(torch.fft.fft2(zs) * wf, norm="forward")
            pl = torch.cat([f[:, :1].real, f[:, :1].imag, f[:, 1:2].real], 1)
            mn = pl.amin(dim=(2, 3), keepdim=True)
            mx = pl.amax(dim=(2, 3), keepdim=True)
            n = (pl - mn) / (mx - mn)
            d_out[b0:b0 + chunk] = torch.clamp(torch.round(n * 255.0), 0, 255).to(torch.uint8).permute(0, 2, 3, 1)

    for _ in range(2):
        cufft_pipeline()
    e0.record(st)
    for _ in range(3):
        cufft_pipeline()
    e1.record(st)
    torch.cuda.synchronize()
    cu = e0.elapsed_time(e1) / 3
    px = B * H * W / 1e6
    print("this repo: %.2f ms (%.0f Mpixel/s) | torch.fft/cuFFT pipeline (same 2-planes-per-transform packing, eager torch ops): %.2f ms (%.0f Mpixel/s)"
          % (ours, px / ours * 1e3, cu, px / cu * 1e3))
except Exception as e:  # torch missing or OOM: the comparison is optional
    print("cuFFT comparison skipped:", e)
