"""Developer diagnostic (not a pytest file): quick numerical sweep on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from conftest import load_fdr, load_oracle, rel_l2, u8_gate
fdr = load_fdr(); O = load_oracle()
print("devices", fdr.device_count())
rng = np.random.default_rng(0)
for n in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]:
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    y = fdr.fft_radix2(x); yi = fdr.fft_radix2(x, True)
    print("fft1d", n, rel_l2(y, np.fft.fft(x.astype(np.complex128))), rel_l2(yi, np.fft.ifft(x.astype(np.complex128)) * n))
for shp in [(1, 8), (8, 1), (4, 4), (16, 32), (64, 128), (128, 64), (512, 1024), (1024, 2048), (2048, 2048), (12, 20), (5, 64)]:
    m = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64)
    y = fdr.dft2d(m); yi = fdr.dft2d(m, True)
    print("dft2d", shp, rel_l2(y, np.fft.fft2(m.astype(np.complex128))), rel_l2(yi, np.fft.ifft2(m.astype(np.complex128)) * m.size))
for S, ang in [(50, 30), (40, 45), (21, 0), (33, -17.5), (64, 90), (7, 360), (1, 10)]:
    with fdr.Plan(128, 128, 1) as p:
        p.set_psf_motion(S, ang)
        a = p.get_psf(); b = O.port().motion_psf(S, ang)
        print("psf", S, ang, "bit-identical" if np.array_equal(a.view(np.uint32), b.view(np.uint32)) else "DIFF %g" % np.abs(a - b).max())
for (H, W, S, ang) in [(48, 80, 9, 30), (64, 64, 5, 10), (100, 200, 21, 45), (330, 640, 40, 45), (7, 9, 3, 20), (1, 33, 1, 0)]:
    planes = [rng.random((H, W), dtype=np.float32) for _ in range(3)]
    psf = O.port().motion_psf(S, ang)
    with fdr.Plan(H, W, 3) as p:
        p.set_psf(psf, 0.01)
        t = time.time(); outs = p.restore_planes(planes); dt = time.time() - t
        Rp, Cp = p.padded
        want_u8, want = O.restore_image_u8(planes, psf, 0.01)
        got_u8 = np.stack([O.port().pack_u8(o) for o in outs], -1)
        res = O.port().wiener_deblur(O.pad_pow2(planes[0]), psf, 0.01, want=("G", "F", "H"))
        G = p.forward_spectrum(planes[0]); F = p.filtered_spectrum(planes[0])
        print("restore", (H, W), "pad", (Rp, Cp), "G", rel_l2(G, res["G"]), "F", rel_l2(F, res["F"]),
              "maxabs", max(np.abs(a - b).max() for a, b in zip(outs, want)), "u8 exact/off1/worse", u8_gate(got_u8, want_u8), "%.1f ms" % (dt * 1e3))
        imgs = np.stack([np.clip(np.rint(pl * 255), 0, 255).astype(np.uint8) for pl in planes], -1)[None]
        out8 = p.restore_images_u8(imgs)
        planes8 = [imgs[0, :, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        w8, _ = O.restore_image_u8(planes8, psf, 0.01)
        print("   u8 path exact/off1/worse", u8_gate(out8[0], w8), "launches", p.last_launch_count(), "profile", p.profile())
