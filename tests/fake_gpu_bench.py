"""TEST HARNESS ONLY -- runs bench.py's control flow on a machine WITHOUT a GPU.

    python -m torch.distributed.run --nproc-per-node 2 ... tests/fake_gpu_bench.py --gpus 2 --steps 2 ...

bench.py's multi-rank line (batch workload + the row-sharded leg under its watchdog) can only be run for real on a multi-GPU
box.  This launcher replaces what needs a device -- torch.cuda streams/events, device tensors, the NCCL process group and the
ctypes binding of libfdr_b200.so -- by CPU stand-ins (numpy float64 pipelines with the same call API), shrinks the workload
geometry, and then calls bench.main() unchanged, so that every Python statement of the N > 1 path (loop counts derived from
all-reduced durations, gathers, the JSON line, the watchdog) executes under gloo exactly as under NCCL.  Nothing here is
reachable from the product or from bench.py itself: the stand-ins are installed by this file only.

Environment: FAKE_TMP (directory shared by the ranks), FAKE_HANG_RANK (that rank never returns from phase 1 of the sharded
leg: exercises the watchdog), FAKE_RAISE_RANK (that rank raises inside the sharded leg), FAKE_CRASH_RANK (that rank's process
kills itself inside the sharded leg), FAKE_CHILD_NO_START (the child processes of the isolated leg exit at once)."""
import contextlib
import ctypes
import os
import sys
import time
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = (ctypes.c_uint8 * n).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


# ---------------------------------------------------------------- torch.cuda / device tensors
class FakeStream:
    _next = [0x1000]

    def __init__(self, device=None, priority=0):
        FakeStream._next[0] += 0x10
        self.cuda_stream = FakeStream._next[0]

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)

    def synchronize(self):
        pass


def install_fake_cuda():
    cur = [FakeStream()]
    tc = torch.cuda
    tc.is_available = lambda: True
    tc.set_device = lambda d: None
    tc.synchronize = lambda *a, **k: None
    tc.empty_cache = lambda: None
    tc.current_stream = lambda device=None: cur[0]
    tc.set_stream = lambda s: cur.__setitem__(0, s)
    tc.Stream = FakeStream
    tc.Event = FakeEvent
    tc.stream = lambda s: contextlib.nullcontext()

    def on_cpu(fn):
        def wrapped(*a, **k):
            dev = k.get("device")
            if dev is not None and torch.device(dev).type == "cuda":
                k["device"] = "cpu"
            k.pop("pin_memory", None)
            return fn(*a, **k)
        return wrapped

    for name in ("empty", "zeros", "tensor", "ones"):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    real_init = dist.init_process_group

    def init_pg(backend=None, **k):
        k.pop("device_id", None)
        return real_init("gloo", **k)

    dist.init_process_group = init_pg


# ---------------------------------------------------------------- stand-in for the ctypes binding (fdr.py)
def _oracle():
    from conftest import load_oracle
    return load_oracle()


def _restore_f64(img_u8, wf, Rp, Cp):
    """(H, W, C) u8 -> (H, W, C) u8, SURVEY.md Appendix A in float64."""
    H, W, C = img_u8.shape
    out = np.empty_like(img_u8)
    for c in range(C):
        g = np.zeros((Rp, Cp))
        g[:H, :W] = img_u8[:, :, c] / 255.0
        f = np.real(np.fft.ifft2(np.fft.fft2(g) * wf))
        n = (f - f.min()) / (f.max() - f.min())
        out[:, :, c] = np.clip(np.rint(n[:H, :W] * 255.0), 0, 255).astype(np.uint8)
    return out


def _wiener(psf, Rp, Cp, K):
    hp = np.zeros((Rp, Cp))
    hp[: psf.shape[0], : psf.shape[1]] = psf
    Hs = np.fft.fft2(hp)
    return np.conj(Hs) / (np.abs(Hs) ** 2 + K)


class FakePinned:
    def __init__(self, shape, dtype):
        self.array = np.zeros(shape, dtype)
        self.nbytes = self.array.nbytes

    def free(self):
        self.array = None


class FakePlan:
    KINDS = ("pass1_rows_fwd", "pass2_cols_wiener", "pass3_rows_inv_minmax", "pass4_normalize_pack")

    def __init__(self, rows, cols, channels=3, max_images=1, device=0):
        self.rows, self.cols, self.channels = rows, cols, channels
        self.padded = (next_pow2(rows), next_pow2(cols))
        self.timing = False
        self.calls = 0

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def set_chunk_images(self, n):
        pass

    def set_psf_motion(self, length, angle, K=0.01):
        self.wf = _wiener(_oracle().port().motion_psf(length, angle), self.padded[0], self.padded[1], K)

    def get_wiener(self):
        return self.wf.astype(np.complex64)

    def _restore(self, a, o):
        for i in range(a.shape[0]):
            o[i] = _restore_f64(a[i], self.wf, *self.padded)
        if self.timing:
            self.calls += 1

    def restore_images_device_u8(self, d_in, d_out, n, stream=0):
        shape = (n, self.rows, self.cols, self.channels)
        self._restore(_arr(d_in, shape, np.uint8), _arr(d_out, shape, np.uint8))

    def restore_images_u8(self, images, out=None):
        if out is None:
            out = np.empty_like(images)
        self._restore(images, out)
        return out

    def set_kernel_timing(self, on=True):
        self.timing = bool(on)

    def kernel_timing(self):
        n = max(1, self.calls)
        return {k: {"ms": 0.1 * (i + 1) * n if k != "pass2_cols_wiener" else 0.9 * n, "launches": n, "bytes": 1e6 * n}
                for i, k in enumerate(self.KINDS)}

    def time_pass(self, which, variant=0, npairs=3, reps=10):
        return 0.05 * which

    def last_launch_count(self):
        return 4


class FakeShard:
    """fdr.Shard's call API on shared-memory files (plane-pair form, exchanges fused into phase 1 / phase 3)."""

    def __init__(self, rows, cols, channels, rank, world, device=0):
        self.rank, self.world, self.C = rank, world, channels
        self.H, self.W = rows, cols
        self.padded_rows, self.padded_cols = next_pow2(rows), next_pow2(cols)
        self.Rl, self.Cl = self.padded_rows // world, self.padded_cols // world
        self.first_row = rank * self.Rl
        self.n_rows = max(0, min(self.first_row + self.Rl, rows) - self.first_row)
        self.cols_per_rank = self.Cl
        self.npairs = (channels + 1) // 2
        self.half_plane = False
        self.staged = False
        self.minmax_negated = False
        self.path = os.path.join(os.environ["FAKE_TMP"], "slab%d_%d.bin" % (rank, int(time.time() * 1e3) % 100000))
        self.shape = (self.npairs, self.padded_rows, self.Cl)
        self.slab = np.memmap(self.path, dtype=np.complex128, mode="w+", shape=self.shape)
        self.slab[:] = 0
        self.slab.flush()
        self.mm = np.zeros((channels, 2), np.float32)
        self.launches = 0

    def close_peers(self):
        self.peers = []

    def close(self):
        pass

    def export_handle(self):
        return self.path

    def set_peers_from_handles(self, handles):
        self.peers = [np.memmap(h, dtype=np.complex128, mode="r+", shape=self.shape) for h in handles]

    def set_minmax_negated(self, on=True):
        self.minmax_negated = bool(on)

    def minmax_ptr(self):
        return self.mm.ctypes.data

    def set_psf_motion(self, length, angle, K=0.01):
        wf = _wiener(_oracle().port().motion_psf(length, angle), self.padded_rows, self.padded_cols, K)
        self.wf = wf[:, self.rank * self.Cl:(self.rank + 1) * self.Cl]

    def phase1(self, d_in_rows, stream=0, pair=None):
        if os.environ.get("FAKE_HANG_RANK") == str(self.rank):
            time.sleep(10 ** 6)
        if os.environ.get("FAKE_CRASH_RANK") == str(self.rank):
            os.kill(os.getpid(), 9)
        if os.environ.get("FAKE_RAISE_RANK") == str(self.rank):
            raise RuntimeError("injected failure on rank %d" % self.rank)
        self.launches = 0
        x = _arr(d_in_rows, (max(self.n_rows, 1), self.W, self.C), np.uint8)[: self.n_rows].astype(np.float64) / 255.0
        for p in range(self.npairs):
            a = x[:, :, 2 * p]
            b = x[:, :, 2 * p + 1] if 2 * p + 1 < self.C else np.zeros_like(a)
            z = np.zeros((self.n_rows, self.padded_cols), np.complex128)
            z[:, : self.W] = a + 1j * b
            Z = np.fft.fft(z, axis=1)
            for g in range(self.world):
                self.peers[g][p, self.first_row:self.first_row + self.n_rows, :] = Z[:, g * self.Cl:(g + 1) * self.Cl]
        for m in self.peers:
            m.flush()
        self.launches += 2

    def exchange1(self, stream=0, pair=None):
        pass

    def exchange3(self, stream=0, pair=None):
        pass

    def phase2(self, stream=0, pair=None):
        slab = np.memmap(self.path, dtype=np.complex128, mode="r+", shape=self.shape)
        for p in range(self.npairs):
            col = np.array(slab[p])
            col[self.H:] = 0   # rows >= H hold the previous image's output: the real column pass zero-fills them (rows_valid)
            slab[p] = np.fft.ifft(np.fft.fft(col, axis=0) * self.wf, axis=0)
        slab.flush()
        self.launches += 3

    def phase3(self, stream=0, pair=None):
        r0 = self.rank * self.Rl
        self.raw = np.zeros((self.C, self.Rl, self.padded_cols))
        for p in range(self.npairs):
            row = np.concatenate([np.array(np.memmap(m.filename, dtype=np.complex128, mode="r", shape=self.shape)[p, r0:r0 + self.Rl, :])
                                  for m in self.peers], axis=1)
            z = np.fft.ifft(row, axis=1)
            self.raw[2 * p] = z.real
            if 2 * p + 1 < self.C:
                self.raw[2 * p + 1] = z.imag
        self.mm[:, 0] = self.raw.min(axis=(1, 2))
        mx = self.raw.max(axis=(1, 2))
        self.mm[:, 1] = -mx if self.minmax_negated else mx
        self.launches += 2

    def phase4(self, d_out_rows, stream=0):
        out = _arr(d_out_rows, (max(self.n_rows, 1), self.W, self.C), np.uint8)
        mn = self.mm[:, 0].astype(np.float64)
        mx = (-self.mm[:, 1] if self.minmax_negated else self.mm[:, 1]).astype(np.float64)
        for c in range(self.C):
            n = (self.raw[c, : self.n_rows, : self.W] - mn[c]) / (mx[c] - mn[c])
            out[: self.n_rows, :, c] = np.clip(np.rint(n * 255.0), 0, 255).astype(np.uint8)
        self.launches += 2

    def sync_timed_out(self, stream=0):
        return False

    def last_launch_count(self):
        return self.launches


def make_fake_fdr():
    m = types.ModuleType("fdr_b200_binding")
    m.Plan = FakePlan
    m.Shard = FakeShard
    m.PinnedArray = FakePinned

    def synth_images_device_u8(d_out, seed, first_image, n, channels, H, W, stream=0):
        out = _arr(d_out, (n, H, W, channels), np.uint8)
        for i in range(n):
            out[i] = np.transpose(_oracle().synth_image_u8(seed - 0xF17E0000, first_image + i, H, W, channels), (1, 2, 0))

    def synth_rows_device_u8(d_out, seed, image, channels, rows_total, cols, first_row, n_rows, stream=0):
        if n_rows == 0:
            return
        img = np.transpose(_oracle().synth_image_u8(seed - 0xF17E0000, image, rows_total, cols, channels), (1, 2, 0))
        _arr(d_out, (n_rows, cols, channels), np.uint8)[:] = img[first_row:first_row + n_rows]

    m.synth_images_device_u8 = synth_images_device_u8
    m.synth_rows_device_u8 = synth_rows_device_u8
    m.l2_flush = lambda ptr, n, stream=0: None
    m.device_count = lambda: 1
    return m


def main():
    if os.environ.get("FAKE_CHILD_NO_START") == "1" and "--sharded-child" in sys.argv:
        return 7   # a box on which the children of the isolated sharded leg cannot even start
    install_fake_cuda()
    import bench
    fake_fdr = make_fake_fdr()
    real_load = bench._load

    def load(name, path):
        if name == "fdr_b200_binding":
            return fake_fdr
        mod = real_load(name, path)
        if name == "fdr_dist":   # its zero-copy device view becomes a view of host memory
            mod.device_tensor = lambda ptr, shape, device, typestr="<f4": torch.from_numpy(_arr(ptr, shape, np.float32))
        return mod

    bench._load = load
    bench.CHILD_ENTRY = os.path.abspath(__file__)   # the children of the isolated sharded leg need the same stand-ins
    small = int(os.environ.get("FAKE_SIZE", "48"))
    bench.WORKLOADS = dict(bench.WORKLOADS)
    bench.WORKLOADS["batch256x2048"] = (3, 2, small, small + 16, 5, 30.0)
    bench.WORKLOADS["rgb16384"] = (4, 1, small + 8, small, 5, 30.0)
    bench.WORKLOADS["rgb4096"] = (2, 1, small, small, 5, 30.0)
    bench.WORKLOADS["car"] = (1, 1, 33, 64, 5, 45.0)
    bench.WORKLOADS["cat"] = (0, 1, 78, 96, 5, 30.0)
    return bench.main()


if __name__ == "__main__":
    sys.exit(main())
