"""CPU tests (no GPU): the C-ABI library loads and exports what include/fdr_b200.h declares, the
host-side argument checks and error convention work, the CLI keeps the reference's contract, and
the compat PNG reader agrees with cv2.  No compute call is made without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, PKG, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(fdr):
    L = fdr.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libfdr_b200.so does not export " + n
    assert L.fdr_version() >= 100


def test_binding_covers_header(fdr):
    """Every declared entry point has ctypes argtypes in the binding (tests call through it)."""
    L = fdr.lib()
    for n in declared_symbols():
        if n in ("fdr_last_error",):
            continue
        assert getattr(L, n).argtypes is not None, n


def test_argument_validation_without_device(fdr):
    L = fdr.lib()
    h = C.c_void_p()
    assert L.fdr_plan_create(C.byref(h), 0, 10, 3, 1, 0) == -1  # FDR_E_INVALID
    assert b"must all be" in L.fdr_last_error()
    assert L.fdr_plan_create(C.byref(h), 20000, 10, 3, 1, 0) == -1
    assert L.fdr_plan_create(None, 8, 8, 3, 1, 0) == -1
    assert L.fdr_plan_destroy(None) == 0
    assert L.fdr_fft_radix2_host(None, 12, 0) == -1  # not a power of two
    assert b"power-of-two" in L.fdr_last_error()
    if fdr.device_count() == 0:
        # no CPU fallback: with no device a well-formed request fails loudly with FDR_E_CUDA
        assert L.fdr_plan_create(C.byref(h), 8, 8, 3, 1, 0) == -2
        x = np.zeros(16, np.float32)
        assert L.fdr_fft_radix2_host(x.ctypes.data_as(C.POINTER(C.c_float)), 8, 0) == -2
        assert len(L.fdr_last_error()) > 0
        with pytest.raises(fdr.FdrError):
            fdr.Plan(8, 8, 3)


def test_cli_contract():
    """gpu.cpp:58-61,69: usage text and exit code -1 on bad argc / unreadable image."""
    exe = os.path.join(PKG, "gpu")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout == "Usage: ./gpu <img-path> <psf-length> <psf-angle>\n"
    r = subprocess.run([exe, "/nonexistent.png", "50", "30"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout == "Cannot read image\n"


def test_compat_png_reader_matches_cv2(tmp_path):
    """compat/cvmat.hpp imread/imwrite (zlib) against cv2 on the reference's sample image."""
    cv2 = pytest.importorskip("cv2")
    src = tmp_path / "t.cpp"
    src.write_text('#define FDR_FORCE_COMPAT_MAT\n#include "compat/cvmat.hpp"\n#include <cstdio>\n'
                   'int main(int c, char** v){ cv::Mat m = cv::imread(v[1]); if (m.empty()) return 2;'
                   ' FILE* f = fopen(v[2], "wb"); fwrite(m.data, 1, m.step * m.rows, f); fclose(f);'
                   ' printf("%d %d\\n", m.rows, m.cols); return cv::imwrite(v[3], m) ? 0 : 3; }\n')
    exe = tmp_path / "t"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", PKG, str(src), "-o", str(exe), "-lz"], check=True, env=env)
    png = os.path.join(GOLDEN, "input", "car_blurred.png")
    raw, out = tmp_path / "raw.bin", tmp_path / "out.png"
    r = subprocess.run([str(exe), png, str(raw), str(out)], capture_output=True, text=True, check=True)
    want = cv2.imread(png, cv2.IMREAD_COLOR)
    assert r.stdout.split() == [str(want.shape[0]), str(want.shape[1])]
    got = np.fromfile(raw, np.uint8).reshape(want.shape)
    assert np.array_equal(got, want)
    assert np.array_equal(cv2.imread(str(out), cv2.IMREAD_COLOR), want)


def test_cpp_layer_defines_the_symbols_the_reference_header_declares(tmp_path):
    """Drop-in at link level: a translation unit that includes the REFERENCE's own fft/fft.hpp
    (fft.hpp:31-45) and takes the address of every fft_gpu:: function must link against this repository's fft/fft_gpu.cpp
    compiled in its `<opencv2/opencv.hpp>` branch (here: the oracle's header stand-in for OpenCV) -- i.e. the mangled names
    and signatures agree, not just the spelling.  Needs /root/reference (build container only)."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "fft", "fft.hpp")):
        pytest.skip("reference sources not present on this box")
    shim = os.path.join(ROOT, "oracle", "cvshim")
    env = dict(os.environ)
    env.pop("CXX", None)
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "fft/fft.hpp"\n'
                     'void* const fdr_probe[] = {(void*)&fft_gpu::wienerDeblur_RGB_naive, (void*)&fft_gpu::wienerDeblur_RGB_optimized,\n'
                     '    (void*)&fft_gpu::fft_radix2_kernel, (void*)&fft_gpu::dft_naive_kernel, (void*)&fft_gpu::transform_row_kernel,\n'
                     '    (void*)&fft_gpu::my_dft2D, (void*)&fft_gpu::wienerDeblur_myfft};\n'
                     'int main() { return fdr_probe[0] == nullptr; }\n')
    subprocess.run(["g++", "-std=c++17", "-O0", "-w", "-I", shim, "-I", ref, "-c", str(probe), "-o", str(tmp_path / "probe.o")],
                   check=True, env=env)
    # this repository's host layer in the OpenCV branch of fft/fft.hpp (no FDR_FORCE_COMPAT_MAT)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", shim, "-I", PKG, "-I", os.path.join(ROOT, "include"), "-c",
                    os.path.join(PKG, "fft", "fft_gpu.cpp"), "-o", str(tmp_path / "fft_gpu.o")], check=True, env=env)

    def syms(obj, flag):
        out = subprocess.run(["nm", flag, str(obj)], capture_output=True, text=True, check=True).stdout
        return {ln.split()[-1] for ln in out.splitlines() if "fft_gpu" in ln}

    wanted = syms(tmp_path / "probe.o", "--undefined-only")
    have = syms(tmp_path / "fft_gpu.o", "--defined-only")
    assert len(wanted) == 7 and wanted <= have, (sorted(wanted - have), sorted(have))
    subprocess.run(["g++", "-o", str(tmp_path / "probe"), str(tmp_path / "probe.o"), str(tmp_path / "fft_gpu.o"),
                    "-L", os.path.join(PKG, "lib"), "-lfdr_b200", "-Wl,-rpath," + os.path.join(PKG, "lib")], check=True, env=env)
