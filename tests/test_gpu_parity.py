"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(include/fdr_b200.h, ctypes binding); the oracle (oracle/, pinned by tests/test_oracle.py) is only
the checker.  Gates are BASELINE.json's: complex spectra within 1e-4 relative L2 (fp32), 8-bit
images within +-1 LSB on >= 99.9 % of pixels, with the exact counts asserted/printed."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, PKG, rel_l2, u8_gate

pytestmark = pytest.mark.gpu

SPECTRUM_TOL = 1e-4      # BASELINE.json north_star: complex spectra, relative L2, fp32
U8_MIN_FRAC = 0.999      # >= 99.9 % of pixels within +-1 LSB
K = 0.01


def check_u8(got, want, allow_worse=0):
    exact, off1, worse = u8_gate(got, want)
    frac = (exact + off1) / got.size
    print("u8 parity: %d px, %d exact, %d off-by-1, %d off-by-more (%.5f%% within 1 LSB)" % (got.size, exact, off1, worse, 100 * frac))
    assert frac >= U8_MIN_FRAC, (exact, off1, worse)
    assert worse <= allow_worse, (exact, off1, worse)
    return exact, off1, worse


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------------------
def test_native_library_is_the_one_running(gpu):
    import ctypes
    assert os.path.exists(gpu.LIB_PATH)
    maps = open("/proc/self/maps").read()
    assert "libfdr_b200.so" in maps
    assert gpu.device_count() >= 1
    assert isinstance(gpu.lib(), ctypes.CDLL)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fft_radix2_all_lengths(gpu, oracle, n):
    """fft_serial.cpp:40-68 against the GPU radix passes, both directions (inverse unscaled)."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for inv in (False, True):
        y = gpu.fft_radix2(x, inv)
        exact = np.fft.ifft(x.astype(np.complex128)) * n if inv else np.fft.fft(x.astype(np.complex128))
        assert rel_l2(y, exact) < 1e-6, (n, inv)
        assert rel_l2(y, oracle.port().fft1d(x, inv)) < SPECTRUM_TOL, (n, inv)


def test_fft_golden_vectors(gpu):
    g = np.load(os.path.join(GOLDEN, "restore_small.npz"))
    assert rel_l2(gpu.fft_radix2(g["fft64_in"]), g["fft64_fwd"]) < SPECTRUM_TOL
    assert rel_l2(gpu.fft_radix2(g["fft64_in"], True), g["fft64_inv"]) < SPECTRUM_TOL
    assert rel_l2(gpu.dft_naive(g["dft12_in"]), g["dft12_fwd"]) < SPECTRUM_TOL
    assert rel_l2(gpu.transform_rows(g["dft12_in"][None, :])[0], g["dft12_fwd"]) < SPECTRUM_TOL
    with pytest.raises(gpu.FdrError):
        gpu.fft_radix2(g["dft12_in"])  # radix-2 entry point rejects non powers of two


@pytest.mark.parametrize("shape", [(1, 1), (1, 8), (8, 1), (4, 4), (16, 32), (128, 64), (512, 1024), (1024, 2048),
                                   (12, 20), (5, 64), (64, 5), (3, 7)])
def test_dft2d_shapes(gpu, oracle, shape):
    """fft_serial.cpp:113-139 my_dft2D incl. non power-of-two lengths (dft_naive path, :71-87)."""
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    m = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    m64 = m.astype(np.complex128)
    assert rel_l2(gpu.dft2d(m), np.fft.fft2(m64)) < 2e-6
    assert rel_l2(gpu.dft2d(m, True), np.fft.ifft2(m64) * m.size) < 2e-6
    if m.size <= 1 << 16:
        assert rel_l2(gpu.dft2d(m), oracle.port().dft2d(m)) < SPECTRUM_TOL


def test_dft2d_roundtrip_4096(gpu):
    """Full-size property: IDFT(DFT(x)) = rows*cols*x (both unscaled, fft_serial.cpp:67)."""
    rng = np.random.default_rng(5)
    m = (rng.standard_normal((4096, 4096)) + 1j * rng.standard_normal((4096, 4096))).astype(np.complex64)
    back = gpu.dft2d(gpu.dft2d(m), True)
    assert rel_l2(back / np.float32(m.size), m) < 1e-6
    # Parseval on the forward spectrum
    F = gpu.dft2d(m)
    e_time = float(np.sum(np.abs(m.astype(np.complex128)) ** 2))
    e_freq = float(np.sum(np.abs(F.astype(np.complex128)) ** 2)) / m.size
    assert abs(e_freq / e_time - 1) < 1e-5


def test_device_psf_bit_exact_with_cv2_golden(gpu):
    """utils.hpp:15-24 built on the device == cv2 4.13 bit for bit."""
    g = np.load(os.path.join(GOLDEN, "psf_cases.npz"))
    for key in g.files:
        _, s, ang = key.split("_")
        got = gpu.motion_psf(int(s), float(ang))
        assert np.array_equal(bits(got), bits(g[key])), key
    with gpu.Plan(64, 64, 1) as p:
        p.set_psf_motion(9, 30.0, K)
        assert np.array_equal(bits(p.get_psf()), bits(g["psf_9_30.0"]))


def test_wiener_factor(gpu, oracle):
    """Wf = conj(H)/(|H|^2+K) (fft_serial.cpp:186-224) built once per plan."""
    psf = oracle.port().motion_psf(21, 45.0)
    with gpu.Plan(100, 200, 1) as p:
        p.set_psf(psf, K)
        Rp, Cp = p.padded
        assert (Rp, Cp) == (128, 256)
        wf = p.get_wiener()
    hp = np.zeros((Rp, Cp))
    hp[:21, :21] = psf
    H = np.fft.fft2(hp)
    want = np.conj(H) / (np.abs(H) ** 2 + np.float32(K))
    assert rel_l2(wf, want) < 2e-6
    assert float(np.abs(wf).max()) <= 5.0 + 1e-3  # |Wf| <= 1/(2 sqrt(K))


RESTORE_CASES = [(48, 80, 9, 30.0), (64, 64, 5, 10.0), (100, 200, 21, 45.0), (7, 9, 3, 20.0), (1, 33, 1, 0.0),
                 (33, 1, 1, 0.0), (1, 1, 1, 0.0), (256, 256, 50, 30.0), (330, 640, 40, 45.0), (17, 300, 15, 123.4),
                 # long columns: K x 2048 blocks (col_blocks.cu) or, for pitches the TMA kernel cannot take, the 128-point four-step (col_split)
                 (8192, 48, 9, 30.0), (5000, 100, 21, 45.0), (16384, 40, 5, 10.0), (9000, 16, 3, 20.0),
                 # 2048 padded rows: the 64-points-per-thread column kernel (col_wide.cu), full and zero-padded columns
                 (2048, 96, 9, 30.0), (1100, 40, 5, 10.0)]


@pytest.mark.parametrize("H,W,S,ang", RESTORE_CASES)
def test_restore_planes_parity(gpu, oracle, H, W, S, ang):
    """fft_gpu::wienerDeblur_RGB_optimized boundary (fft.hpp:33) vs the serial oracle: spectra,
    normalised planes, 8-bit pack.  Includes ragged / 1-pixel / non-pow2 sizes."""
    rng = np.random.default_rng(H * 1000 + W)
    planes = [rng.random((H, W), dtype=np.float32) for _ in range(3)]
    psf = oracle.port().motion_psf(S, ang)
    with gpu.Plan(H, W, 3) as p:
        p.set_psf(psf, K)
        outs = p.restore_planes(planes)
        G = p.forward_spectrum(planes[0])
        F = p.filtered_spectrum(planes[0])
        mm = p.last_minmax(3)
    res = oracle.port().wiener_deblur(oracle.pad_pow2(planes[0]), psf, K, want=("G", "F", "norm", "raw"))
    assert rel_l2(G, res["G"]) < SPECTRUM_TOL
    assert rel_l2(F, res["F"]) < SPECTRUM_TOL
    want_u8, want = oracle.restore_image_u8(planes, psf, K)
    lo, hi = res["minmax"]
    if hi - lo > 1e-3 * max(abs(hi), 1e-30):  # degenerate (1x1) planes have no range
        assert abs(mm[0, 0] - lo) <= 1e-4 * (hi - lo) and abs(mm[0, 1] - hi) <= 1e-4 * (hi - lo)
        for a, b in zip(outs, want):
            assert np.abs(a - b).max() < 1e-4
    got_u8 = np.stack([oracle.port().pack_u8(o) for o in outs], -1)
    if H * W >= 1000:
        check_u8(got_u8, want_u8)
    else:
        assert np.abs(got_u8.astype(int) - want_u8.astype(int)).max() <= 1


def test_restore_golden_reference_outputs(gpu):
    """Committed outputs of the reference's own compiled serial code (restore_small.npz)."""
    g = np.load(os.path.join(GOLDEN, "restore_small.npz"))
    for name in ("a", "b"):
        img, psf = g[name + "_img"], g[name + "_psf"]
        with gpu.Plan(img.shape[0], img.shape[1], 1) as p:
            p.set_psf(psf, K)
            out = p.restore_planes([img])[0]
            G = p.forward_spectrum(img)
        assert rel_l2(G, g[name + "_G"]) < SPECTRUM_TOL
        assert np.abs(out - g[name + "_norm"][: img.shape[0], : img.shape[1]]).max() < 1e-4


def test_wiener_factor_long_columns(gpu, oracle):
    """Plans with 8192/16384 rows keep Wf in digit-swapped row order; the inspection call hands it
    back in natural order."""
    psf = oracle.port().motion_psf(9, 30.0)
    with gpu.Plan(8192, 32, 1) as p:
        p.set_psf(psf, K)
        wf = p.get_wiener()
    hp = np.zeros((8192, 32))
    hp[:9, :9] = psf
    H = np.fft.fft2(hp)
    assert rel_l2(wf, np.conj(H) / (np.abs(H) ** 2 + np.float32(K))) < 2e-6


def test_errors(gpu, oracle):
    with gpu.Plan(16, 16, 3) as p:
        with pytest.raises(gpu.FdrError):  # restore before a PSF is set
            p.restore_planes([np.zeros((16, 16), np.float32)])
        with pytest.raises(gpu.FdrError):  # PSF larger than the padded image (copyMakeBorder would throw)
            p.set_psf(np.ones((40, 40), np.float32), K)
        p.set_psf_motion(5, 10.0, K)
        with pytest.raises(gpu.FdrError):  # 8-bit output needs whole images
            gpu._check(gpu.lib().fdr_restore_planes_device_f32(p.h, 1, None, 1, 4, None))


@pytest.mark.parametrize("name", ["car", "cat"])
def test_sample_images(gpu, oracle, name):
    """BASELINE configs[0], [1]: the reference's sample images, whole-image u8 path (imread bytes
    in, restored bytes out) against the reference-serial output pinned by sample_hashes.json."""
    cv2 = pytest.importorskip("cv2")
    h = json.load(open(os.path.join(GOLDEN, "sample_hashes.json")))[name]
    bgr = cv2.imread(os.path.join(GOLDEN, "input", name + "_blurred.png"), cv2.IMREAD_COLOR)
    H, W, _ = bgr.shape
    planes = [bgr[:, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
    psf = oracle.port().motion_psf(*h["psf"])
    want, outs = oracle.restore_image_u8(planes, psf, K)
    assert hashlib.sha256(want.tobytes()).hexdigest() == h["sha256_u8"]  # oracle == reference bytes
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(h["psf"][0], h["psf"][1], K)
        got = p.restore_images_u8(bgr[None])[0]
        got_planes = p.restore_planes(planes)
    check_u8(got, want)
    for c in range(3):
        assert abs(float(got_planes[c].mean(dtype=np.float64)) - h["mean"][c]) < 1e-5
        assert np.abs(got_planes[c] - outs[c]).max() < 1e-4


def test_cli_end_to_end(gpu, oracle, tmp_path):
    """./gpu <img> <len> <angle> [out.png]: the reference CLI contract (gpu.cpp:57-138) + output."""
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(PKG, "gpu")
    png = os.path.join(GOLDEN, "input", "car_blurred.png")
    out = tmp_path / "restored.png"
    r = subprocess.run([exe, png, "40", "45", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for needle in ("=== FAST (Reuse Memory) Profiling (3 Channels) ===", "=== SLOW (Naive Allocation) Profiling (3 Channels) ===",
                   "Deblurring 3 channels took(gpu[optimize]): ", "Deblurring 3 channels took(gpu): ",
                   "[1. Allocation]  Time: ", "[4. GPU Compute] Time: ", "Total (Sum)      Time: "):
        assert needle in r.stdout, needle
    got = cv2.imread(str(out), cv2.IMREAD_COLOR)  # content is checked in tests/test_gpu_whitebalance.py
    assert got is not None and got.shape == (330, 640, 3)


def test_batch_chunking_and_pairing(gpu, oracle):
    """Images are packed two planes per complex transform and processed in chunks; neither may
    change the bytes.  Odd plane counts leave a half-empty pair."""
    H, W, n = 96, 160, 5
    imgs = np.stack([np.transpose(oracle.synth_image_u8(7, i, H, W), (1, 2, 0)) for i in range(n)])
    psf = oracle.port().motion_psf(9, 30.0)
    outs = []
    for chunk in (0, 1, 2, 3, 5):
        with gpu.Plan(H, W, 3, max_images=n) as p:
            p.set_psf(psf, K)
            p.set_chunk_images(chunk)
            outs.append(p.restore_images_u8(imgs))
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    for i in (0, n - 1):
        planes = [imgs[i, :, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        want, _ = oracle.restore_image_u8(planes, psf, K)
        check_u8(outs[0][i], want)
    # planes entry point with 1, 2, 7 planes == per-plane results
    rng = np.random.default_rng(0)
    pls = [rng.random((H, W), dtype=np.float32) for _ in range(7)]
    with gpu.Plan(H, W, 1) as p:
        p.set_psf(psf, K)
        all7 = p.restore_planes(pls)
        for k in (0, 3, 6):
            one = p.restore_planes([pls[k]])[0]
            assert np.abs(one - all7[k]).max() < 2e-6  # pairing partner differs -> fp32 cross-talk only
            want = oracle.port().wiener_deblur(oracle.pad_pow2(pls[k]), psf, K)["norm"][:H, :W]
            assert np.abs(all7[k] - want).max() < 1e-4


def test_linearity_property_2048(gpu, oracle):
    """Size-independent property at a BASELINE size: the un-normalised restoration is linear, so
    min/max of restore(c*x) scale by c and the normalised output is unchanged."""
    H = W = 2048
    img = oracle.synth_image_u8(3, 0, H, W, channels=1)[0].astype(np.float32) * np.float32(1.0 / 255.0)
    with gpu.Plan(H, W, 1) as p:
        p.set_psf_motion(50, 30.0, K)
        a = p.restore_planes([img])[0]
        mm_a = p.last_minmax(1)[0]
        b = p.restore_planes([img * np.float32(0.5)])[0]
        mm_b = p.last_minmax(1)[0]
    assert np.allclose(mm_b, 0.5 * mm_a, rtol=1e-5)
    assert np.abs(a - b).max() < 1e-5


def test_full_size_2048_batch_vs_serial(gpu, oracle):
    """BASELINE configs[3] shape (reduced count): 6 synthetic 2048x2048x3 images, first and last
    checked against the serial oracle, the rest against per-image runs of the same code."""
    H = W = 2048
    n = 6
    imgs = np.stack([np.transpose(oracle.synth_image_u8(3, i, H, W), (1, 2, 0)) for i in range(n)])
    with gpu.Plan(H, W, 3, max_images=n) as p:
        p.set_psf_motion(50, 30.0, K)
        out = p.restore_images_u8(imgs)
        assert p.last_launch_count() > 0
        single = p.restore_images_u8(imgs[2:3])[0]
    # a plane's pairing partner differs between the two calls: fp32 cross-talk only
    ex, off1, worse = u8_gate(single, out[2])
    assert worse == 0 and off1 <= 1e-4 * single.size, (ex, off1, worse)
    psf = oracle.port().motion_psf(50, 30.0)
    for i in (0, n - 1):
        planes = [imgs[i, :, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        want, _ = oracle.restore_image_u8(planes, psf, K)
        check_u8(out[i], want)


def test_full_size_4096_vs_serial(gpu, oracle):
    """BASELINE configs[2]: synthetic 4096x4096, one channel against the serial oracle (~10 s CPU),
    spectra gate included."""
    H = W = 4096
    img8 = oracle.synth_image_u8(2, 0, H, W)
    images = np.ascontiguousarray(np.transpose(img8, (1, 2, 0)))[None]
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(50, 30.0, K)
        out = p.restore_images_u8(images)[0]
        plane = img8[1].astype(np.float32) * np.float32(1.0 / 255.0)
        G = p.forward_spectrum(plane)
        F = p.filtered_spectrum(plane)
    res = oracle.port().wiener_deblur(plane, oracle.port().motion_psf(50, 30.0), K, want=("G", "F", "norm"))
    assert rel_l2(G, res["G"]) < SPECTRUM_TOL
    assert rel_l2(F, res["F"]) < SPECTRUM_TOL
    check_u8(out[:, :, 1], oracle.port().pack_u8(res["norm"]))


def test_long_transform_16384_separable(gpu):
    """N = 16384 rows and columns (BASELINE configs[4] lengths) on one GPU: the 2-D spectrum of a
    separable plane a[y]*b[x] is the outer product of two 1-D spectra (float64 oracle)."""
    n = 16384
    rng = np.random.default_rng(11)
    a = rng.random(n).astype(np.float32)
    b = rng.random(n).astype(np.float32)
    plane = np.outer(a, b).astype(np.float32)
    with gpu.Plan(n, n, 1) as p:
        G = p.forward_spectrum(plane)
    A = np.fft.fft(a.astype(np.float64))
    B = np.fft.fft(b.astype(np.float64))
    rows = rng.integers(0, n, 64)
    for r in rows:
        want = A[r] * B
        assert rel_l2(G[r], want) < 5e-6 * max(1.0, np.abs(A[0]) / max(np.abs(A[r]), 1e-9)) + 2e-6 or rel_l2(G[r], want) < SPECTRUM_TOL


def test_kernel_timing_api(gpu, oracle):
    H, W = 256, 256
    imgs = np.transpose(oracle.synth_image_u8(7, 0, H, W), (1, 2, 0))[None]
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(9, 30.0, K)
        p.set_kernel_timing(True)
        p.restore_images_u8(imgs)
        kt = p.kernel_timing()
        assert p.last_launch_count() == 4  # rows, columns, rows, pack (slot reset and fold ride along in pass 1 and the pack)
    assert all(kt[k]["launches"] == 1 and kt[k]["ms"] > 0 and kt[k]["bytes"] > 0 for k in kt)


HALF_CASES = [(48, 80, 9, 30.0), (64, 64, 5, 10.0), (100, 200, 21, 45.0), (7, 70, 3, 20.0), (2, 64, 1, 0.0), (3, 33, 1, 0.0),
              (256, 256, 50, 30.0), (330, 640, 40, 45.0), (17, 300, 15, 123.4), (782, 1920, 50, 30.0),
              (8192, 96, 9, 30.0), (5000, 100, 21, 45.0), (16384, 80, 5, 10.0), (9000, 130, 3, 20.0),
              (2048, 96, 9, 30.0), (1100, 130, 5, 10.0), (4096, 200, 9, 30.0), (1024, 2048, 9, 30.0)]


@pytest.mark.parametrize("H,W,S,ang", HALF_CASES)
def test_half_plane_lone_plane_parity(gpu, oracle, monkeypatch, H, W, S, ang):
    """An odd plane count sends the last plane through the half-plane path (two rows of the plane per complex row
    transform, Hermitian half spectrum + Nyquist column; passes.h).  Forced on for every size here
    (FDR_HALF_MIN_PIXELS=0); gates are the oracle's, for 3 planes (one pair + the lone plane) and for 1 plane."""
    monkeypatch.setenv("FDR_HALF_MIN_PIXELS", "0")
    monkeypatch.setenv("FDR_HALF", "1")
    rng = np.random.default_rng(H * 1000 + W + 1)
    planes = [rng.random((H, W), dtype=np.float32) for _ in range(3)]
    psf = oracle.port().motion_psf(S, ang)
    with gpu.Plan(H, W, 3) as p:
        p.set_psf(psf, K)
        expect_half = p.padded[1] >= 64 and p.padded[0] >= 2
        assert p.half_plane == expect_half
        outs = p.restore_planes(planes)
        mm = p.last_minmax(3)
        lone_only = p.restore_planes([planes[2]])[0]
    want_u8, want = oracle.restore_image_u8(planes, psf, K)
    res = oracle.port().wiener_deblur(oracle.pad_pow2(planes[2]), psf, K, want=("norm", "raw"))
    lo, hi = res["minmax"]
    assert abs(mm[2, 0] - lo) <= 1e-4 * (hi - lo) and abs(mm[2, 1] - hi) <= 1e-4 * (hi - lo)
    for a, b in zip(outs, want):
        assert np.abs(a - b).max() < 1e-4
    assert np.abs(lone_only - want[2]).max() < 1e-4
    got_u8 = np.stack([oracle.port().pack_u8(o) for o in outs], -1)
    if H * W >= 1000:
        check_u8(got_u8, want_u8)
    else:
        assert np.abs(got_u8.astype(int) - want_u8.astype(int)).max() <= 1


def test_half_plane_u8_images_and_chunks(gpu, oracle, monkeypatch):
    """u8 image entry point with the half-plane path forced on: one image per chunk (3 planes: a pair + the lone plane)
    against two images per chunk (6 planes: pairs only) -- the two arithmetic paths agree to <= 1 LSB -- and the oracle."""
    monkeypatch.setenv("FDR_HALF_MIN_PIXELS", "0")
    H, W, n = 200, 300, 4
    imgs = np.stack([np.transpose(oracle.synth_image_u8(7, i, H, W), (1, 2, 0)) for i in range(n)])
    psf = oracle.port().motion_psf(9, 30.0)
    outs = []
    for chunk in (1, 2):
        with gpu.Plan(H, W, 3, max_images=n) as p:
            p.set_psf(psf, K)
            p.set_chunk_images(chunk)
            assert p.half_plane
            outs.append(p.restore_images_u8(imgs))
    ex, off1, worse = u8_gate(outs[0], outs[1])
    assert worse == 0 and off1 <= 2e-3 * outs[0].size, (ex, off1, worse)
    for i in (0, n - 1):
        planes = [imgs[i, :, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        want, _ = oracle.restore_image_u8(planes, psf, K)
        check_u8(outs[0][i], want)


def test_plan_calls_on_two_streams_are_ordered(gpu, oracle):
    """One workspace per plan: device restores issued on two different streams (and a PSF rebuild in between) are ordered
    through the plan's event instead of racing (include/fdr_b200.h, "streams")."""
    torch = pytest.importorskip("torch")
    H = W = 1024
    dev = torch.device("cuda", 0)
    img = np.ascontiguousarray(np.transpose(oracle.synth_image_u8(7, 1, H, W), (1, 2, 0)))
    d_in = torch.from_numpy(img).to(dev)
    outs = [torch.zeros_like(d_in) for _ in range(3)]
    sts = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(9, 30.0, K)
        p.restore_images_device_u8(d_in.data_ptr(), outs[0].data_ptr(), 1, sts[0].cuda_stream)
        p.restore_images_device_u8(d_in.data_ptr(), outs[1].data_ptr(), 1, sts[1].cuda_stream)
        p.set_psf_motion(9, 30.0, K)  # rebuilds the factor through the shared workspace
        p.restore_images_device_u8(d_in.data_ptr(), outs[2].data_ptr(), 1, sts[0].cuda_stream)
        torch.cuda.synchronize()
    a, b, c = (o.cpu().numpy() for o in outs)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    planes = [img[:, :, k].astype(np.float32) * np.float32(1.0 / 255.0) for k in range(3)]
    want, _ = oracle.restore_image_u8(planes, oracle.port().motion_psf(9, 30.0), K)
    check_u8(a, want)


def test_many_tiny_images_grid_cap(gpu, oracle):
    """More plane pairs than gridDim.y allows in one launch: the chunk size is capped (65534 planes)."""
    H, W, n = 4, 4, 44000  # 132000 planes
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    psf = np.zeros((3, 3), np.float32)
    psf[1, :] = 1.0 / 3
    with gpu.Plan(H, W, 3, max_images=n) as p:
        p.set_psf(psf, K)
        out = p.restore_images_u8(imgs)
    for i in (0, 21844, 21845, n - 1):
        planes = [imgs[i, :, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
        want, _ = oracle.restore_image_u8(planes, psf, K)
        assert np.abs(out[i].astype(int) - want.astype(int)).max() <= 1


def test_full_size_16384_against_serial_oracle(gpu, oracle):
    """BASELINE configs[4] at full size, one colour plane against the serial restatement of fft_serial.cpp:141-261 (the
    long-row kernels, the K x 2048 column blocks and the half-plane path all at their real geometry): forward and filtered
    spectra within 1e-4 relative L2, 8-bit plane within +-1 LSB on >= 99.9 % of the pixels, exact counts printed.  About two
    minutes of serial CPU work."""
    H = W = 16384
    img = oracle.synth_image_u8(4, 0, H, W)                 # (3, H, W)
    psf = oracle.port().motion_psf(50, 30.0)
    plane0 = img[0].astype(np.float32) * np.float32(1.0 / 255.0)
    res = oracle.port().wiener_deblur(plane0, psf, K, want=("G", "F", "norm"))
    want_u8 = oracle.port().pack_u8(res["norm"])
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(50, 30.0, K)
        G = p.forward_spectrum(plane0)
        eg = rel_l2(G, res["G"])
        del G
        F = p.filtered_spectrum(plane0)
        ef = rel_l2(F, res["F"])
        del F
        got = p.restore_images_u8(np.ascontiguousarray(np.transpose(img, (1, 2, 0)))[None])[0]
        half = p.half_plane
    print("16384^2: forward spectrum rel-L2 %.3g, filtered spectrum rel-L2 %.3g (gate %g), half-plane path for plane 2: %s"
          % (eg, ef, SPECTRUM_TOL, half))
    assert eg < SPECTRUM_TOL and ef < SPECTRUM_TOL
    check_u8(got[:, :, 0], want_u8)
    # planes 1 and 2 (plane 2 takes the half-plane path) against the reference's own openmp build when it is there
    if oracle.have_ref():
        oracle.ref().set_threads(os.cpu_count() or 1)
        for c in (1, 2):
            norm = oracle.ref().wiener(img[c].astype(np.float32) * np.float32(1.0 / 255.0), psf, K, "openmp")
            check_u8(got[:, :, c], oracle.port().pack_u8(norm))
