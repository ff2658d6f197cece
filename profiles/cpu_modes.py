"""The reference's CPU modes timed on THIS host: serial, openmp, simd compiled
unmodified into oracle/_ref/libref.so, and the MPI mode over the single-node stand-in
(oracle/_ref/ref_mpi_bench).  Wall clock around the 3-channel loop, as the reference drivers measure it
(serial.cpp:33-41, openmp.cpp:89-100, mpi.cpp:95-111).
Usage: python profiles/cpu_modes.py > profiles/r1/cpu_modes.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_oracle  # noqa: E402

O = load_oracle()
R = O.ref()
cores = os.cpu_count() or 1
print("host cores: %d" % cores)
cases = {"car 330x640 (pad 512x1024), psf 40/45": (1, 330, 640, 40, 45.0), "cat 782x1920 (pad 1024x2048), psf 50/30": (0, 782, 1920, 50, 30.0),
         "synthetic 2048x2048, psf 50/30": (3, 2048, 2048, 50, 30.0), "synthetic 4096x4096, psf 50/30": (2, 4096, 4096, 50, 30.0)}
devnull = os.open(os.devnull, os.O_WRONLY)
for name, (cfg, H, W, S, ang) in cases.items():
    img = O.synth_image_u8(cfg, 0, H, W)
    planes = np.stack([O.pad_pow2(img[c].astype(np.float32) * np.float32(1.0 / 255.0)) for c in range(3)])
    psf = O.port().motion_psf(S, ang)
    res = {}
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # the reference prints per-stage timers
    try:
        for mode, thr in (("serial", 1), ("simd", 1), ("openmp", cores), ("openmp", 12)):
            if mode == "openmp":
                R.set_threads(thr)
                R.wiener(planes[0], psf, 0.01, mode)  # thread-pool warm-up (openmp.cpp runs serial first)
            t0 = time.perf_counter()
            for pl in planes:
                R.wiener(pl, psf, 0.01, mode)
            res["%s (%d thread%s)" % (mode, thr, "s" if thr > 1 else "")] = (time.perf_counter() - t0) * 1e3
        if O.have_ref_mpi():
            for ranks in sorted({4, min(cores, 16)}):
                res["mpi stand-in (%d ranks)" % ranks] = O.ref_mpi_wiener(planes, psf, 0.01, ranks)[1]
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
    print(name)
    for k, v in res.items():
        print("   %-28s %10.1f ms   %8.3f Mpixel/s" % (k, v, H * W / v / 1e3))
