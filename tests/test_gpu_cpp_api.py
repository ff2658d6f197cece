"""GPU test of the C++ host layer: every fft_gpu:: function of fft/fft.hpp (fft.hpp:31-45), compiled
from tests/cpp/api_check.cpp against fft/fft_gpu.cpp + libfdr_b200.so, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, rel_l2

pytestmark = pytest.mark.gpu


def test_cpp_interface(gpu, oracle, tmp_path):
    exe = tmp_path / "api_check"
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    subprocess.run(["g++", "-std=c++17", "-O1", "-DFDR_FORCE_COMPAT_MAT", "-I", PKG, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "api_check.cpp"), os.path.join(PKG, "fft", "fft_gpu.cpp"),
                    "-L", os.path.join(PKG, "lib"), "-lfdr_b200", "-Wl,-rpath," + os.path.join(PKG, "lib"), "-lz", "-o", str(exe)],
                   check=True, env=env)
    out = tmp_path / "dump.bin"
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "=== FAST (Reuse Memory) Profiling (3 Channels) ===" in r.stdout and "=== SLOW (Naive Allocation)" in r.stdout
    d = np.fromfile(out, np.float32)
    pos = 0

    def take(n):
        nonlocal pos
        v = d[pos:pos + n]
        pos += n
        return v

    def cplx(v):
        return v.view(np.complex64)

    N, R, C, H, W = 64, 16, 32, 40, 56
    a, b, m = cplx(take(2 * N)), cplx(take(24)), cplx(take(2 * R * C)).reshape(R, C)
    P = oracle.port()
    assert rel_l2(cplx(take(2 * N)), P.fft1d(a)) < 1e-4
    assert rel_l2(cplx(take(2 * N)), P.fft1d(a, True)) < 1e-4
    assert rel_l2(cplx(take(24)), P.fft1d(b)) < 1e-4
    assert rel_l2(cplx(take(24)), P.fft1d(b)) < 1e-4
    fwd = cplx(take(2 * R * C)).reshape(R, C)
    assert rel_l2(fwd, P.dft2d(m)) < 1e-4
    back = cplx(take(2 * R * C)).reshape(R, C)
    assert rel_l2(back, m * (R * C)) < 1e-5
    img = take(H * W).reshape(H, W).copy()
    psf = take(81).reshape(9, 9).copy()
    assert np.array_equal(psf.view(np.uint32), P.motion_psf(9, 30.0).view(np.uint32))
    want = P.wiener_deblur(oracle.pad_pow2(img), psf, 0.01)["norm"][:H, :W]
    assert np.abs(take(H * W).reshape(H, W) - want).max() < 1e-4
    for _ in range(5):  # 3 planes of RGB_optimized (one was a strided ROI) + 2 planes of RGB_naive
        assert np.abs(take(H * W).reshape(H, W) - want).max() < 1e-4
    assert pos == d.size
