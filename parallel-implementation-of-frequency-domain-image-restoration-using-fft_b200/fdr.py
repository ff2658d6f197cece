"""ctypes binding of lib/libfdr_b200.so (the C ABI in include/fdr_b200.h).

Test / benchmark harness only: the product's host side is the C++ in fft/ and gpu.cpp.
The binding never computes anything itself and has NO fallback: if the shared library is
missing or a call fails, it raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfdr_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "fdr_b200.h")

_fp = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


class FdrError(RuntimeError):
    pass


def build(jobs=8, quiet=True):
    """Compile the CUDA library (and the ./gpu CLI) for sm_100a with the package Makefile."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", HERE, "-j%d" % jobs, "all"], check=True, env=env,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:  # the in-tree library normally travels with the snapshot; build it if it did not
                build()
            except Exception as e:
                raise FdrError("CUDA library not built and building failed: %s (run `make -C %s`): %s" % (LIB_PATH, HERE, e))
        L = C.CDLL(LIB_PATH)
        L.fdr_last_error.restype = C.c_char_p
        vp, i, f, d, sz, ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_longlong
        pp = C.POINTER(C.c_void_p)
        sig = {
            "fdr_version": [],
            "fdr_device_count": [C.POINTER(i)],
            "fdr_host_alloc": [pp, sz],
            "fdr_host_free": [vp],
            "fdr_plan_create": [pp, i, i, i, i, i],
            "fdr_plan_destroy": [vp],
            "fdr_plan_padded_size": [vp, C.POINTER(i), C.POINTER(i)],
            "fdr_plan_set_chunk_images": [vp, i],
            "fdr_plan_set_white_balance": [vp, i],
            "fdr_plan_half_plane": [vp, C.POINTER(i)],
            "fdr_white_balance_pack_host": [C.POINTER(_fp), C.POINTER(_fp), i, i, vp],
            "fdr_plan_set_psf_host": [vp, _fp, i, i, sz, f],
            "fdr_plan_set_psf_motion": [vp, i, d, f],
            "fdr_plan_get_psf_host": [vp, _fp, i, C.POINTER(i), C.POINTER(i)],
            "fdr_plan_get_wiener_host": [vp, _fp],
            "fdr_restore_planes_host_f32": [vp, C.POINTER(_fp), sz, C.POINTER(_fp), sz, i],
            "fdr_restore_images_host_u8": [vp, vp, vp, i],
            "fdr_restore_images_device_u8": [vp, vp, vp, i, vp],
            "fdr_restore_planes_device_f32": [vp, vp, vp, vp, i, vp],
            "fdr_plan_last_minmax_host": [vp, _fp, i],
            "fdr_plan_get_profile": [vp, _fp],
            "fdr_plan_last_launch_count": [vp, C.POINTER(ll)],
            "fdr_plan_time_pass": [vp, i, i, i, i, _fp],
            "fdr_plan_set_kernel_timing": [vp, i],
            "fdr_plan_get_kernel_timing": [vp, C.POINTER(d), C.POINTER(ll), C.POINTER(d)],
            "fdr_plan_forward_spectrum_host": [vp, _fp, sz, _fp],
            "fdr_plan_filtered_spectrum_host": [vp, _fp, sz, _fp],
            "fdr_dft2d_host": [_fp, i, i, i],
            "fdr_fft_radix2_host": [_fp, i, i],
            "fdr_dft_naive_host": [_fp, i, i],
            "fdr_transform_rows_host": [_fp, i, i, i],
            "fdr_motion_psf_host": [i, d, _fp],
            "fdr_memcpy": [vp, vp, sz, i],
            "fdr_shard_create": [pp, i, i, i, i, i, i],
            "fdr_shard_destroy": [vp],
            "fdr_shard_geometry": [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)],
            "fdr_shard_local_slab": [vp, pp, C.POINTER(sz)],
            "fdr_ipc_export": [vp, C.c_char_p],
            "fdr_ipc_open": [C.c_char_p, pp],
            "fdr_ipc_close": [vp],
            "fdr_shard_set_peers": [vp, pp],
            "fdr_shard_set_psf_motion": [vp, i, d, f],
            "fdr_shard_set_psf_host": [vp, _fp, i, i, f],
            "fdr_shard_phase1_rows": [vp, vp, vp],
            "fdr_shard_phase2_cols": [vp, vp],
            "fdr_shard_phase3_rows": [vp, vp],
            "fdr_shard_phase1_pairs": [vp, vp, i, i, vp],
            "fdr_shard_phase2_pairs": [vp, i, i, vp],
            "fdr_shard_phase3_pairs": [vp, i, i, vp],
            "fdr_shard_pair_count": [vp, C.POINTER(i)],
            "fdr_shard_half_plane": [vp, C.POINTER(i)],
            "fdr_shard_set_row_ctas": [vp, i],
            "fdr_shard_set_minmax_negated": [vp, i],
            "fdr_shard_exchange1": [vp, i, i, vp],
            "fdr_shard_exchange3": [vp, i, i, vp],
            "fdr_shard_staged": [vp, C.POINTER(i)],
            "fdr_shard_set_link_ctas": [vp, i],
            "fdr_shard_set_ce_streams": [vp, i],
            "fdr_shard_restore_rows": [vp, vp, vp, vp],
            "fdr_shard_timeline": [vp, _fp, i, C.POINTER(i)],
            "fdr_shard_barrier": [vp, i, vp],
            "fdr_shard_minmax_allreduce": [vp, vp],
            "fdr_shard_sync_status": [vp, vp, C.POINTER(i)],
            "fdr_shard_minmax_device": [vp, pp],
            "fdr_shard_phase4_pack": [vp, vp, vp],
            "fdr_shard_last_launch_count": [vp, C.POINTER(ll)],
            "fdr_synth_rows_device_u8": [vp, C.c_uint32, ll, i, i, i, i, i, vp],
            "fdr_synth_images_device_u8": [vp, C.c_uint32, ll, i, i, i, i, vp],
            "fdr_l2_flush_device": [vp, sz, vp],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def exported_symbols():
    return [s for s in dir(lib()) if s.startswith("fdr_")]


def _check(rc):
    if rc != 0:
        raise FdrError("fdr error %d: %s" % (rc, lib().fdr_last_error().decode()))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


def device_count():
    n = C.c_int(0)
    rc = lib().fdr_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class PinnedArray:
    """numpy view over cudaMallocHost memory."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = C.c_void_p()
        _check(lib().fdr_host_alloc(C.byref(self.ptr), self.nbytes))
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().fdr_host_free(self.ptr)
            self.ptr = None


class Plan:
    """One image geometry (rows x cols x channels) + one PSF/K: mirrors what
    fft_gpu::wienerDeblur_RGB_optimized sets up per call (fft_gpu.cu:279-322)."""

    def __init__(self, rows, cols, channels=3, max_images=1, device=0):
        self.h = C.c_void_p()
        self.rows, self.cols, self.channels = rows, cols, channels
        _check(lib().fdr_plan_create(C.byref(self.h), rows, cols, channels, max_images, device))
        pr, pc = C.c_int(), C.c_int()
        _check(lib().fdr_plan_padded_size(self.h, C.byref(pr), C.byref(pc)))
        self.padded = (pr.value, pc.value)

    def close(self):
        if self.h:
            lib().fdr_plan_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_chunk_images(self, n):
        _check(lib().fdr_plan_set_chunk_images(self.h, n))

    def set_white_balance(self, on=True):
        _check(lib().fdr_plan_set_white_balance(self.h, int(on)))

    @property
    def half_plane(self):
        v = C.c_int(0)
        _check(lib().fdr_plan_half_plane(self.h, C.byref(v)))
        return bool(v.value)

    def set_psf(self, psf, K=0.01):
        psf = _f32(psf)
        _check(lib().fdr_plan_set_psf_host(self.h, _p(psf), psf.shape[0], psf.shape[1], 0, K))

    def set_psf_motion(self, length, angle, K=0.01):
        _check(lib().fdr_plan_set_psf_motion(self.h, int(length), float(angle), K))

    def get_psf(self):
        r, c = C.c_int(), C.c_int()
        _check(lib().fdr_plan_get_psf_host(self.h, None, 0, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32)
        _check(lib().fdr_plan_get_psf_host(self.h, _p(out), out.size, C.byref(r), C.byref(c)))
        return out

    def get_wiener(self):
        out = np.empty(self.padded, np.complex64)
        _check(lib().fdr_plan_get_wiener_host(self.h, _p(out.view(np.float32))))
        return out

    def restore_planes(self, planes):
        """list of (rows, cols) f32 planes -> list of normalised f32 planes (host API)."""
        ins = [_f32(p) for p in planes]
        outs = [np.empty((self.rows, self.cols), np.float32) for _ in ins]
        n = len(ins)
        ip = (_fp * n)(*[_p(a) for a in ins])
        op = (_fp * n)(*[_p(a) for a in outs])
        _check(lib().fdr_restore_planes_host_f32(self.h, ip, 0, op, 0, n))
        return outs

    def restore_images_u8(self, images, out=None):
        """u8 (n, rows, cols, channels) -> u8 restored images (host API).  `images`/`out` may be
        PinnedArray.array views to skip the staging copy."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        n = images.shape[0]
        assert images.shape[1:] == (self.rows, self.cols, self.channels), images.shape
        if out is None:
            out = np.empty_like(images)
        _check(lib().fdr_restore_images_host_u8(self.h, images.ctypes.data, out.ctypes.data, n))
        return out

    def restore_images_device_u8(self, d_in, d_out, n_images, stream=0):
        _check(lib().fdr_restore_images_device_u8(self.h, d_in, d_out, n_images, stream))

    def restore_planes_device_f32(self, d_in, d_out_f32, d_out_u8, n_planes, stream=0):
        _check(lib().fdr_restore_planes_device_f32(self.h, d_in, d_out_f32, d_out_u8, n_planes, stream))

    def last_minmax(self, n_planes):
        out = np.zeros((n_planes, 2), np.float32)
        _check(lib().fdr_plan_last_minmax_host(self.h, _p(out), n_planes))
        return out

    def profile(self):
        out = np.zeros(6, np.float32)
        _check(lib().fdr_plan_get_profile(self.h, _p(out)))
        return out

    def last_launch_count(self):
        n = C.c_longlong(0)
        _check(lib().fdr_plan_last_launch_count(self.h, C.byref(n)))
        return n.value

    def time_pass(self, which, variant=0, npairs=3, reps=10):
        ms = C.c_float(0)
        _check(lib().fdr_plan_time_pass(self.h, which, variant, npairs, reps, C.byref(ms)))
        return ms.value

    def set_kernel_timing(self, on=True):
        _check(lib().fdr_plan_set_kernel_timing(self.h, int(on)))

    KERNEL_KINDS = ("pass1_rows_fwd", "pass2_cols_wiener", "pass3_rows_inv_minmax", "pass4_normalize_pack")

    def kernel_timing(self):
        ms = (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        b = (C.c_double * 4)()
        _check(lib().fdr_plan_get_kernel_timing(self.h, ms, n, b))
        return {k: {"ms": ms[i], "launches": n[i], "bytes": b[i]} for i, k in enumerate(self.KERNEL_KINDS)}

    def forward_spectrum(self, plane):
        plane = _f32(plane)
        out = np.empty(self.padded, np.complex64)
        _check(lib().fdr_plan_forward_spectrum_host(self.h, _p(plane), 0, _p(out.view(np.float32))))
        return out

    def filtered_spectrum(self, plane):
        plane = _f32(plane)
        out = np.empty(self.padded, np.complex64)
        _check(lib().fdr_plan_filtered_spectrum_host(self.h, _p(plane), 0, _p(out.view(np.float32))))
        return out


class Shard:
    """One rank's share of a row-sharded restoration (include/fdr_b200.h, fdr_shard_*)."""

    def __init__(self, rows, cols, channels, rank, world, device=0):
        self.h = C.c_void_p()
        self.rows, self.cols, self.channels, self.rank, self.world = rows, cols, channels, rank, world
        _check(lib().fdr_shard_create(C.byref(self.h), rows, cols, channels, rank, world, device))
        v = [C.c_int() for _ in range(5)]
        _check(lib().fdr_shard_geometry(self.h, *[C.byref(x) for x in v]))
        self.first_row, self.n_rows, self.padded_rows, self.padded_cols, self.cols_per_rank = [x.value for x in v]
        self._opened = []

    def close_peers(self):
        """Unmap the peers' slabs (fdr_ipc_close).  With several processes, every rank must have done this BEFORE any rank
        frees its slab (close()): CUDA leaves freeing exported memory that a peer still has mapped undefined -- put a
        cross-rank barrier between the two."""
        for ptr in self._opened:
            lib().fdr_ipc_close(ptr)
        self._opened = []

    def close(self):
        if self.h:
            self.close_peers()
            lib().fdr_shard_destroy(self.h)
            self.h = None

    def local_slab(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(lib().fdr_shard_local_slab(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def export_handle(self):
        buf = C.create_string_buffer(64)
        _check(lib().fdr_ipc_export(self.local_slab()[0], buf))
        return buf.raw

    def set_peers_from_handles(self, handles):
        ptrs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.local_slab()[0])
            else:
                ptr = C.c_void_p()
                _check(lib().fdr_ipc_open(hb, C.byref(ptr)))
                self._opened.append(ptr)
                ptrs.append(ptr.value)
        self.set_peers(ptrs)

    def set_peers(self, ptrs):
        arr = (C.c_void_p * self.world)(*ptrs)
        _check(lib().fdr_shard_set_peers(self.h, arr))

    def set_psf_motion(self, length, angle, K=0.01):
        _check(lib().fdr_shard_set_psf_motion(self.h, int(length), float(angle), K))

    def set_psf(self, psf, K=0.01):
        psf = _f32(psf)
        _check(lib().fdr_shard_set_psf_host(self.h, _p(psf), psf.shape[0], psf.shape[1], K))

    @property
    def npairs(self):
        n = C.c_int(0)
        _check(lib().fdr_shard_pair_count(self.h, C.byref(n)))
        return n.value

    @property
    def half_plane(self):
        v = C.c_int(0)
        _check(lib().fdr_shard_half_plane(self.h, C.byref(v)))
        return bool(v.value)

    def set_row_ctas(self, n):
        _check(lib().fdr_shard_set_row_ctas(self.h, int(n)))

    def set_minmax_negated(self, on=True):
        _check(lib().fdr_shard_set_minmax_negated(self.h, int(on)))
        self.minmax_negated = bool(on)

    @property
    def staged(self):
        v = C.c_int(0)
        _check(lib().fdr_shard_staged(self.h, C.byref(v)))
        return bool(v.value)

    def set_link_ctas(self, n):
        _check(lib().fdr_shard_set_link_ctas(self.h, int(n)))

    def set_ce_streams(self, n):
        _check(lib().fdr_shard_set_ce_streams(self.h, int(n)))

    def exchange1(self, stream=0, pair=None):
        first, count = (0, self.npairs) if pair is None else (pair, 1)
        _check(lib().fdr_shard_exchange1(self.h, first, count, stream))

    def exchange3(self, stream=0, pair=None):
        first, count = (0, self.npairs) if pair is None else (pair, 1)
        _check(lib().fdr_shard_exchange3(self.h, first, count, stream))

    def restore_rows_native(self, d_in_rows, d_out_rows, stream=0):
        _check(lib().fdr_shard_restore_rows(self.h, d_in_rows, d_out_rows, stream))

    def timeline(self):
        """FDR_SHARD_TIMELINE=1: {step: [ms from the start of the last native restore to the end of the step, per unit]}."""
        buf = np.zeros(128, np.float32)
        n = C.c_int(0)
        _check(lib().fdr_shard_timeline(self.h, _p(buf), 128, C.byref(n)))
        U = (n.value - 1) // 7
        names = ["phase1", "exchange1", "barrier1", "phase2", "exchange3", "barrier3", "phase3"]
        out = {nm: [round(float(x), 4) for x in buf[k * U:(k + 1) * U]] for k, nm in enumerate(names)}
        out["phase4"] = round(float(buf[7 * U]), 4)
        return out

    def peer_barrier(self, set_index, stream=0):
        _check(lib().fdr_shard_barrier(self.h, int(set_index), stream))

    def minmax_allreduce(self, stream=0):
        _check(lib().fdr_shard_minmax_allreduce(self.h, stream))

    def sync_timed_out(self, stream=0):
        v = C.c_int(0)
        _check(lib().fdr_shard_sync_status(self.h, stream, C.byref(v)))
        return bool(v.value)

    def phase1(self, d_in_rows, stream=0, pair=None):
        if pair is None:
            _check(lib().fdr_shard_phase1_rows(self.h, d_in_rows, stream))
        else:
            _check(lib().fdr_shard_phase1_pairs(self.h, d_in_rows, pair, 1, stream))

    def phase2(self, stream=0, pair=None):
        if pair is None:
            _check(lib().fdr_shard_phase2_cols(self.h, stream))
        else:
            _check(lib().fdr_shard_phase2_pairs(self.h, pair, 1, stream))

    def phase3(self, stream=0, pair=None):
        if pair is None:
            _check(lib().fdr_shard_phase3_rows(self.h, stream))
        else:
            _check(lib().fdr_shard_phase3_pairs(self.h, pair, 1, stream))

    def minmax_ptr(self):
        ptr = C.c_void_p()
        _check(lib().fdr_shard_minmax_device(self.h, C.byref(ptr)))
        return ptr.value

    def phase4(self, d_out_rows, stream=0):
        _check(lib().fdr_shard_phase4_pack(self.h, d_out_rows, stream))

    def last_launch_count(self):
        n = C.c_longlong(0)
        _check(lib().fdr_shard_last_launch_count(self.h, C.byref(n)))
        return n.value


def white_balance_pack(restored_planes, original_planes):
    """gpu.cpp:123-134 on the device: 3 restored + 3 original f32 planes (B, G, R) -> u8 (rows, cols, 3)."""
    r = [_f32(p) for p in restored_planes]
    o = [_f32(p) for p in original_planes]
    rows, cols = r[0].shape
    out = np.empty((rows, cols, 3), np.uint8)
    rp = (_fp * 3)(*[_p(a) for a in r])
    op = (_fp * 3)(*[_p(a) for a in o])
    _check(lib().fdr_white_balance_pack_host(rp, op, rows, cols, out.ctypes.data))
    return out


def memcpy(dst, src, nbytes, kind):
    """kind: 0 = H2D, 1 = D2H, 2 = D2D (raw pointers / numpy .ctypes.data)."""
    _check(lib().fdr_memcpy(dst, src, nbytes, kind))


def synth_rows_device_u8(d_out, seed, image, channels, rows_total, cols, first_row, n_rows, stream=0):
    _check(lib().fdr_synth_rows_device_u8(d_out, seed, image, channels, rows_total, cols, first_row, n_rows, stream))


def dft2d(m, inverse=False):
    a = np.ascontiguousarray(m, dtype=np.complex64).copy()
    _check(lib().fdr_dft2d_host(_p(a.view(np.float32)), a.shape[0], a.shape[1], int(inverse)))
    return a


def fft_radix2(x, inverse=False):
    a = np.ascontiguousarray(x, dtype=np.complex64).copy()
    _check(lib().fdr_fft_radix2_host(_p(a.view(np.float32)), a.shape[0], int(inverse)))
    return a


def dft_naive(x, inverse=False):
    a = np.ascontiguousarray(x, dtype=np.complex64).copy()
    _check(lib().fdr_dft_naive_host(_p(a.view(np.float32)), a.shape[0], int(inverse)))
    return a


def transform_rows(m, inverse=False):
    a = np.ascontiguousarray(m, dtype=np.complex64).copy()
    _check(lib().fdr_transform_rows_host(_p(a.view(np.float32)), a.shape[0], a.shape[1], int(inverse)))
    return a


def motion_psf(length, angle):
    """utils.hpp:15-24 motionBlurKernel, built on the device."""
    out = np.empty((length, length), np.float32)
    _check(lib().fdr_motion_psf_host(int(length), float(angle), _p(out)))
    return out


def synth_images_device_u8(d_out, seed, first_image, n_images, channels, rows, cols, stream=0):
    _check(lib().fdr_synth_images_device_u8(d_out, seed, first_image, n_images, channels, rows, cols, stream))


def l2_flush(d_scratch, nbytes, stream=0):
    _check(lib().fdr_l2_flush_device(d_scratch, nbytes, stream))
