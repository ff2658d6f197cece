"""One-GPU emulation of the 8-way row-sharded 16384^2 restoration (8 slabs of 2048 columns in one process)
against the unsharded plan: debugging aid for the sharded path at its real geometry."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from conftest import load_fdr
import test_gpu_sharded as T
fdr = load_fdr()
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
img = np.random.default_rng(1).integers(0, 256, (H, W, 3), dtype=np.uint8)
got, launches = T.run_emulated(fdr, torch, img, world, 50, 30.0)
with fdr.Plan(H, W, 3) as p:
    p.set_psf_motion(50, 30.0, 0.01)
    single = p.restore_images_u8(img[None])[0]
d = np.abs(got.astype(np.int16) - single.astype(np.int16))
print("emulated world=%d %dx%d vs single: exact %d off1 %d more %d" % (world, H, W, int((d == 0).sum()), int((d == 1).sum()), int((d > 1).sum())))
for r in range(world):
    rows = slice(r * H // world, (r + 1) * H // world)
    print(" rows of rank", r, "max|d|", int(d[rows].max()), "nonzero", int((d[rows] > 0).sum()))
