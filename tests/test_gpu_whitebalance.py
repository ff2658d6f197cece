"""GPU tests of the drivers' post stage: Lab white balance + 8-bit pack
(gpu.cpp:123-134, utils.hpp:55-71) on the device, against cv2 (the library the reference calls).
OpenCV interpolates a fixed-point table for float Lab; the device evaluates the closed forms, so the
gate is the 8-bit one: |delta| <= 1 LSB on >= 99.9 % of pixels, none above 1."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, PKG, u8_gate

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
K = 0.01


def cv2_post_stage(restored_planes, bgr_u8):
    """serial.cpp:43-54 with cv2: merge -> BGR2Lab -> applyWhiteBalance -> Lab2BGR -> convertTo(CV_8U, 255)."""
    img = bgr_u8.astype(np.float32) * np.float32(1.0 / 255.0)
    merged = np.ascontiguousarray(np.stack(restored_planes, -1).astype(np.float32))
    lab = cv2.cvtColor(merged, cv2.COLOR_BGR2Lab)
    lab_o = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
    gain = float(cv2.mean(lab_o[:, :, 0])[0]) / (float(cv2.mean(lab[:, :, 0])[0]) + 1e-6)
    L = np.minimum(np.maximum(lab[:, :, 0] * np.float32(gain), 0.0), 100.0).astype(np.float32)
    lab2 = np.ascontiguousarray(np.stack([L, lab[:, :, 1], lab[:, :, 2]], -1))
    out = cv2.cvtColor(lab2, cv2.COLOR_Lab2BGR)
    return np.clip(np.rint(out * np.float32(255.0)), 0, 255).astype(np.uint8), gain


@pytest.mark.parametrize("name", ["car", "cat"])
def test_white_balance_pipeline_on_samples(gpu, oracle, name):
    psf_par = {"car": (40, 45.0), "cat": (50, 30.0)}[name]
    bgr = cv2.imread(os.path.join(GOLDEN, "input", name + "_blurred.png"), cv2.IMREAD_COLOR)
    H, W, _ = bgr.shape
    planes = [bgr[:, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
    _, restored = oracle.restore_image_u8(planes, oracle.port().motion_psf(*psf_par), K)
    want, gain = cv2_post_stage(restored, bgr)
    assert abs(gain - {"car": 1.0872, "cat": 1.7463}[name]) < 2e-3  # SURVEY.md 8(c) sanity values
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(psf_par[0], psf_par[1], K)
        p.set_white_balance(True)
        got = p.restore_images_u8(bgr[None])[0]
    exact, off1, worse = u8_gate(got, want)
    print("white balance %s: %d exact, %d off by 1, %d off by more" % (name, exact, off1, worse))
    assert worse == 0 and (exact + off1) / got.size >= 0.999
    # the stand-alone post stage on the oracle's planes
    got2 = gpu.white_balance_pack(restored, planes)
    exact, off1, worse = u8_gate(got2, want)
    assert worse == 0 and (exact + off1) / got2.size >= 0.999


def test_cli_writes_white_balanced_image(gpu, oracle, tmp_path):
    exe = os.path.join(PKG, "gpu")
    png = os.path.join(GOLDEN, "input", "car_blurred.png")
    out = tmp_path / "wb.png"
    r = subprocess.run([exe, png, "40", "45", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    bgr = cv2.imread(png, cv2.IMREAD_COLOR)
    planes = [bgr[:, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
    _, restored = oracle.restore_image_u8(planes, oracle.port().motion_psf(40, 45.0), K)
    want, _ = cv2_post_stage(restored, bgr)
    got = cv2.imread(str(out), cv2.IMREAD_COLOR)
    exact, off1, worse = u8_gate(got, want)
    assert worse == 0 and (exact + off1) / got.size >= 0.999
