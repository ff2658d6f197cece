"""Side comparison, run as a subprocess of bench.py (the reference's CHECK_CUDA calls exit() on any error, fft_gpu.cu:59-66,
so it must not share a process with the measurement): the reference's own gpu mode -- fft/fft_gpu.cu compiled unmodified for
sm_100a (oracle/_ref/libref_gpu.so) -- on one HxWx3 image through its 3-plane host boundary, wall clock as gpu.cpp:96-105
takes it, and this library through the same boundary.  Prints one JSON line.
Usage: python profiles/side_refgpu.py H W psf_len psf_angle"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_fdr, load_oracle  # noqa: E402

H, W, S, ang = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
K = 0.01
fdr = load_fdr()
O = load_oracle()
psf = O.port().motion_psf(S, ang)
planes = np.stack([O.synth_image_u8(3, 0, H, W)[c].astype(np.float32) / np.float32(255) for c in range(3)])
fpp = C.POINTER(C.c_float)
out = {}
with fdr.Plan(H, W, 1) as p1:
    p1.set_psf(psf, K)
    p1.restore_planes(list(planes))
    to = []
    for _ in range(3):
        t0 = time.perf_counter()
        p1.restore_planes(list(planes))
        to.append((time.perf_counter() - t0) * 1e3)
out["this_library_same_boundary_ms"] = min(to)
real_stdout = os.dup(1)
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so"))
lib.ref_gpu_restore.restype = C.c_double
lib.ref_gpu_restore.argtypes = [C.c_int, fpp, C.c_int, C.c_int, C.c_int, fpp, C.c_int, C.c_int, C.c_float]
devnull = os.open(os.devnull, os.O_WRONLY)
sys.stdout.flush()
os.dup2(devnull, 1)  # the reference prints its profile block on every call
buf = planes.copy()
lib.ref_gpu_restore(0, buf.ctypes.data_as(fpp), 3, H, W, psf.ctypes.data_as(fpp), S, S, K)  # warm-up, as gpu.cpp:96
ts = []
for _ in range(3):
    buf = planes.copy()
    ts.append(lib.ref_gpu_restore(0, buf.ctypes.data_as(fpp), 3, H, W, psf.ctypes.data_as(fpp), S, S, K))
sys.stdout.flush()
os.dup2(real_stdout, 1)
out.update({"ms_per_image": min(ts), "Mpixel/s": H * W / (min(ts) * 1e-3) / 1e6, "speedup": min(ts) / min(to),
            "what": "fft_gpu::wienerDeblur_RGB_optimized of the reference (fft_gpu.cu:279-394, unmodified, sm_100a) on one %dx%dx3 "
                    "image, 3 f32 host planes in and out, wall clock as gpu.cpp:96-105" % (H, W)})
print(json.dumps(out))
