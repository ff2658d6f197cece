"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

ctypes access to the two CPU checkers:

* ``port``  -- oracle/liboracle.so, the plain-C restatement (wiener_oracle.c) of the
  reference's serial path (/root/reference/fft/fft_serial.cpp, utils.hpp).
* ``ref``   -- oracle/_ref/libref.so, the reference's own serial / openmp / simd
  translation units compiled unmodified (oracle/Makefile, ref_driver.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing here touches the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")

MODE_SERIAL, MODE_OPENMP, MODE_SIMD = 0, 1, 2
_MODES = {"serial": 0, "openmp": 1, "simd": 2}

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def build(quiet=True):
    """Compile liboracle.so and, where /root/reference exists, _ref/libref.so."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", HERE], check=True, env=env,
                   stdout=subprocess.DEVNULL if quiet else None)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=_fp):
    return a.ctypes.data_as(t) if a is not None else t()


class _Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.orc_fft_radix2.argtypes = [_fp, C.c_int, C.c_int]
        L.orc_dft_naive.argtypes = [_fp, C.c_int, C.c_int]
        L.orc_dft2d.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
        L.orc_wiener.argtypes = [_fp, _fp, C.c_size_t, C.c_float]
        L.orc_normalize_minmax.argtypes = [_fp, C.c_size_t, _dp]
        L.orc_wiener_deblur.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_float,
                                        _fp, _fp, _fp, _fp, _fp, _dp]
        L.orc_wiener_deblur.restype = C.c_int
        L.orc_motion_psf.argtypes = [C.c_int, C.c_double, _fp]
        L.orc_pack_u8.argtypes = [_fp, C.c_size_t, _u8p]
        L.orc_synth_u8.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, _u8p]
        L.orc_next_pow2.argtypes = [C.c_int]
        L.orc_next_pow2.restype = C.c_int
        self.L = L

    def fft1d(self, x, inverse=False):
        """x: complex64 (n,) -> complex64 (n,)   [fft_serial.cpp:40-68]"""
        a = np.ascontiguousarray(x, dtype=np.complex64).copy()
        n = a.shape[0]
        if n & (n - 1) == 0:
            self.L.orc_fft_radix2(_ptr(a.view(np.float32)), n, int(inverse))
        else:
            self.L.orc_dft_naive(_ptr(a.view(np.float32)), n, int(inverse))
        return a

    def dft2d(self, m, inverse=False):
        """m: complex64 (rows, cols) -> complex64   [fft_serial.cpp:113-139]"""
        a = np.ascontiguousarray(m, dtype=np.complex64).copy()
        self.L.orc_dft2d(_ptr(a.view(np.float32)), a.shape[0], a.shape[1], int(inverse))
        return a

    def wiener_deblur(self, img, psf, K=0.01, want=("norm",)):
        """img: f32 (Rp, Cp) pow2-padded plane; psf: f32 (S, S).
        Returns dict with the requested keys of norm, G, H, F, raw, minmax
        [fft_serial.cpp:141-261]."""
        img = _f32(img)
        psf = _f32(psf)
        r, c = img.shape
        out = {}
        bufs = {}
        for key, shape, dt in (("norm", (r, c), np.float32), ("G", (r, c), np.complex64),
                               ("H", (r, c), np.complex64), ("F", (r, c), np.complex64),
                               ("raw", (r, c), np.float32)):
            bufs[key] = np.empty(shape, dt) if key in want else None
        mm = np.zeros(2, np.float64)

        def p(key):
            b = bufs[key]
            if b is None:
                return _fp()
            return _ptr(b.view(np.float32) if b.dtype == np.complex64 else b)

        rc = self.L.orc_wiener_deblur(_ptr(img), r, c, _ptr(psf), psf.shape[0], psf.shape[1],
                                      C.c_float(K), p("norm"), p("G"), p("H"), p("F"), p("raw"),
                                      mm.ctypes.data_as(_dp))
        if rc != 0:
            raise RuntimeError("orc_wiener_deblur rc=%d" % rc)
        for k, b in bufs.items():
            if b is not None:
                out[k] = b
        out["minmax"] = (float(mm[0]), float(mm[1]))
        return out

    def motion_psf(self, size, angle):
        """utils.hpp:15-24 with OpenCV's fixed-point bilinear warp."""
        out = np.empty((size, size), np.float32)
        self.L.orc_motion_psf(int(size), float(angle), _ptr(out))
        return out

    def pack_u8(self, x):
        x = _f32(x)
        out = np.empty(x.shape, np.uint8)
        self.L.orc_pack_u8(_ptr(x), x.size, _ptr(out, _u8p))
        return out

    def synth_u8(self, seed, idx0, count):
        out = np.empty(int(count), np.uint8)
        self.L.orc_synth_u8(C.c_uint32(seed), C.c_uint64(idx0), C.c_uint64(count), _ptr(out, _u8p))
        return out

    def next_pow2(self, n):
        return int(self.L.orc_next_pow2(int(n)))


class _Ref:
    def __init__(self):
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (needs /root/reference at build time)")
        L = C.CDLL(REF_SO)
        L.ref_fft1d.argtypes = [C.c_int, _fp, C.c_int, C.c_int]
        L.ref_dft_naive.argtypes = [_fp, C.c_int, C.c_int]
        L.ref_dft2d.argtypes = [C.c_int, _fp, C.c_int, C.c_int, C.c_int]
        L.ref_wiener.argtypes = [C.c_int, _fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_float, _fp]
        L.ref_wiener.restype = C.c_int
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        self.L = L

    def fft1d(self, x, inverse=False, mode="serial"):
        a = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.L.ref_fft1d(_MODES[mode], _ptr(a.view(np.float32)), a.shape[0], int(inverse))
        return a

    def dft_naive(self, x, inverse=False):
        a = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.L.ref_dft_naive(_ptr(a.view(np.float32)), a.shape[0], int(inverse))
        return a

    def dft2d(self, m, inverse=False, mode="serial"):
        a = np.ascontiguousarray(m, dtype=np.complex64).copy()
        self.L.ref_dft2d(_MODES[mode], _ptr(a.view(np.float32)), a.shape[0], a.shape[1], int(inverse))
        return a

    def wiener(self, img, psf, K=0.01, mode="serial"):
        """The reference's wienerDeblur_myfft on a pow2-padded plane -> normalised f32 plane."""
        img = _f32(img)
        psf = _f32(psf)
        out = np.empty_like(img)
        rc = self.L.ref_wiener(_MODES[mode], _ptr(img), img.shape[0], img.shape[1], _ptr(psf),
                               psf.shape[0], psf.shape[1], C.c_float(K), _ptr(out))
        if rc != 0:
            raise RuntimeError("ref_wiener rc=%d" % rc)
        return out

    def set_threads(self, n):
        self.L.ref_set_threads(int(n))

    def max_threads(self):
        return int(self.L.ref_max_threads())


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


def have_ref():
    return os.path.exists(REF_SO)


REF_MPI_BIN = os.path.join(HERE, "_ref", "ref_mpi_bench")


def have_ref_mpi():
    return os.path.exists(REF_MPI_BIN)


def ref_mpi_wiener(planes, psf, K=0.01, nprocs=2):
    """The reference's MPI backend (fft_mpi.cpp compiled unmodified over oracle/mpi_standin) on `nprocs`
    forked ranks, driven like mpi.cpp:95-111.  planes: (n, Rp, Cp) f32, already padded to powers of two.
    Returns (normalised planes, wall ms of the channel loop on rank 0)."""
    import tempfile
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    psf = _f32(psf)
    n, r, c = planes.shape
    with tempfile.TemporaryDirectory() as d:
        pin, ppsf, pout = (os.path.join(d, x) for x in ("in.f32", "psf.f32", "out.f32"))
        planes.tofile(pin)
        psf.tofile(ppsf)
        res = subprocess.run([REF_MPI_BIN, str(int(nprocs)), pin, str(n), str(r), str(c), ppsf, str(psf.shape[0]), repr(float(K)), pout],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("ref_mpi_bench failed: " + res.stderr[-500:])
        ms = [float(ln.split()[1]) for ln in res.stdout.splitlines() if ln.startswith("mpi_ms")][-1]
        out = np.fromfile(pout, np.float32).reshape(n, r, c)
    return out, ms


# ---- helpers shared by tests and bench (pure numpy, follow the reference drivers) ----

def pad_pow2(plane):
    """utils.hpp:40-47 autoPadToPowerOfTwo: zero-pad bottom/right."""
    plane = _f32(plane)
    h, w = plane.shape
    rp, cp = 1, 1
    while rp < h:
        rp <<= 1
    while cp < w:
        cp <<= 1
    out = np.zeros((rp, cp), np.float32)
    out[:h, :w] = plane
    return out


def restore_image_u8(planes_f32, psf, K=0.01, impl=None):
    """serial.cpp:33-39 per-channel loop (pad -> wienerDeblur_myfft -> crop) followed by the
    direct 8-bit pack (others/fft_image_restoration_opencv.cpp:84-86).  planes_f32: list of
    (H, W) f32 planes in [0,1].  Returns (u8 HxWxC, [normalised cropped f32 planes])."""
    impl = impl or (lambda p, k: port().wiener_deblur(p, k, K)["norm"])
    outs = []
    for pl in planes_f32:
        h, w = pl.shape
        outs.append(impl(pad_pow2(pl), psf)[:h, :w].copy())
    u8 = np.stack([port().pack_u8(o) for o in outs], axis=-1)
    return u8, outs


def synth_image_u8(config_index, img, H, W, channels=3):
    """idx = ((img*3+c)*H + y)*W + x ; seed = 0xF17E0000 + config_index.
    Returns u8 (channels, H, W) planes (B, G, R order)."""
    seed = 0xF17E0000 + config_index
    out = np.empty((channels, H, W), np.uint8)
    for c in range(channels):
        idx0 = (img * 3 + c) * H * W
        out[c] = port().synth_u8(seed, idx0, H * W).reshape(H, W)
    return out
