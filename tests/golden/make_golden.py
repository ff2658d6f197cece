"""Regenerates the committed golden vectors.  Run in the BUILD container (needs cv2 and
/root/reference so that oracle/_ref/libref.so holds the reference's own compiled code):

    python tests/golden/make_golden.py

What is pinned, and by whom:
  psf_cases.npz        motionBlurKernel(S, angle) -- cv2 4.13 getRotationMatrix2D + warpAffine,
                       i.e. the third-party arithmetic the reference delegates to OpenCV
                       (utils.hpp:15-24).
  normalize_case.npz   cv2.normalize(NORM_MINMAX, 0, 1) on a random plane (fft_serial.cpp:246).
  restore_small.npz    fft_serial::wienerDeblur_myfft + my_dft2D_forward run from the
                       UNMODIFIED reference sources (oracle/_ref) on two small planes.
  sample_hashes.json   sha256 of the reference-serial restored 8-bit car / cat images
                       (direct x255 pack) and channel statistics (SURVEY.md 8c sanity values).
  input/*.png          the reference's two sample inputs (data fixtures, not source).
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

PSF_CASES = [(50, 30.0), (40, 45.0), (21, 0.0), (33, -17.5), (64, 90.0), (15, 123.4), (7, 360.0),
             (101, 12.25), (2, 45.0), (1, 10.0), (9, 30.0), (5, 10.0), (3, 20.0)]


def cv_psf(S, ang):
    k = np.zeros((S, S), np.float32)
    k[S // 2, :] = np.float32(1.0 / S)
    M = cv2.getRotationMatrix2D((S // 2, S // 2), ang, 1)
    return cv2.warpAffine(k, M, (S, S))


def main():
    ref = O.ref()
    out = {}
    for S, ang in PSF_CASES:
        out["psf_%d_%s" % (S, repr(ang))] = cv_psf(S, ang)
    np.savez_compressed(os.path.join(HERE, "psf_cases.npz"), **out)

    rng = np.random.default_rng(1234)
    x = (rng.standard_normal((64, 96)) * 37.0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "normalize_case.npz"), x=x, y=cv2.normalize(x, None, 0, 1, cv2.NORM_MINMAX))

    small = {}
    for name, (H, W, S, ang, cfg) in {"a": (48, 80, 9, 30.0, 11), "b": (30, 20, 5, 10.0, 12)}.items():
        img = O.synth_image_u8(cfg, 0, H, W, channels=1)[0].astype(np.float32) * np.float32(1.0 / 255.0)
        psf = cv_psf(S, ang)
        padded = O.pad_pow2(img)
        small[name + "_img"] = img
        small[name + "_psf"] = psf
        small[name + "_norm"] = ref.wiener(padded, psf, 0.01, "serial")
        small[name + "_G"] = ref.dft2d(padded.astype(np.complex64), False, "serial")
    xs = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    small["fft64_in"] = xs
    small["fft64_fwd"] = ref.fft1d(xs, False)
    small["fft64_inv"] = ref.fft1d(xs, True)
    xs = (rng.standard_normal(12) + 1j * rng.standard_normal(12)).astype(np.complex64)
    small["dft12_in"] = xs
    small["dft12_fwd"] = ref.dft_naive(xs, False)
    np.savez_compressed(os.path.join(HERE, "restore_small.npz"), **small)

    hashes = {}
    for name, (S, ang) in {"car": (40, 45.0), "cat": (50, 30.0)}.items():
        bgr = cv2.imread(os.path.join(HERE, "input", name + "_blurred.png"), cv2.IMREAD_COLOR)
        planes = [bgr[:, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        psf = cv_psf(S, ang)
        u8, outs = O.restore_image_u8(planes, psf, 0.01, impl=lambda p, k: ref.wiener(p, k, 0.01, "serial"))
        hashes[name] = {
            "psf": [S, ang], "shape": list(u8.shape), "sha256_u8": hashlib.sha256(u8.tobytes()).hexdigest(),
            "sha256_input_bgr": hashlib.sha256(bgr.tobytes()).hexdigest(),
            "mean": [float(o.mean(dtype=np.float64)) for o in outs],
            "std": [float(o.std(dtype=np.float64)) for o in outs],
            "min": [float(o.min()) for o in outs], "max": [float(o.max()) for o in outs],
        }
        print(name, hashes[name])
    json.dump(hashes, open(os.path.join(HERE, "sample_hashes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
