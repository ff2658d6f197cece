"""Kernel-level timing probe (CUDA events inside the library).
python profiles/time_passes.py N [npairs ...]   -- per-pass device time on npairs plane pairs of N x N."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_fdr
fdr = load_fdr()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
pairs = [int(x) for x in sys.argv[2:]] or [max(1, min(12, (1 << 28) // (n * n * 8)))]
with fdr.Plan(n, n, 3) as p:
    p.set_psf_motion(50, 30.0, 0.01)
    for npairs in pairs:
        px = n * n * npairs
        def gbs(ms, bpp): return px * bpp / (ms * 1e-3) / 1e9
        r1 = p.time_pass(1, 0, npairs); r3 = p.time_pass(3, 0, npairs)
        c0 = p.time_pass(2, 0, npairs); c1 = p.time_pass(2, 1, npairs); c2 = p.time_pass(2, 2, npairs); c3 = p.time_pass(2, 3, npairs)
        print("N=%d pairs=%d | per pair: pass1 %.1f us (%.0f GB/s) | pass3 %.1f us (%.0f GB/s) | pass2 default %.1f us (%.0f GB/s), plain-load kernel %.1f us, single FFT %.1f us, copy-only %.1f us (%.0f GB/s)"
              % (n, npairs, r1 * 1e3 / npairs, gbs(r1, 10), r3 * 1e3 / npairs, gbs(r3, 16), c0 * 1e3 / npairs, gbs(c0, 24), c1 * 1e3 / npairs, c2 * 1e3 / npairs, c3 * 1e3 / npairs, gbs(c3, 16)))
