"""CPU tests (gloo, world_size 2): the partitioner's host logic -- row-slab math, handle exchange,
stream-ordered barriers, min/max all-reduce -- driven through fdr_dist.ShardedRestorer with a numpy
stand-in for the CUDA shard that uses the SAME index formulas as the kernels
(peer = x >> log2(Cp/world), slab row = global row)."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT, _load

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def test_partition_helpers():
    d = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
    for H in (1, 5, 330, 782, 1024, 16384):
        for world in (1, 2, 4, 8):
            if d.next_pow2(H) < world:
                continue
            slabs = [d.row_slab(r, world, H) for r in range(world)]
            assert sum(n for _, n in slabs) == H
            pos = 0
            for first, n in slabs:
                if n:
                    assert first == pos
                    pos += n
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            blocks = [d.images_for_rank(r, world, n) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
            assert blocks[0][0] == 0 and all(blocks[i + 1][0] == blocks[i][0] + blocks[i][1] for i in range(world - 1))


class NumpyShard:
    """Test double with the fdr.Shard phase API.  Slabs live in a shared-memory file per rank so the
    'peer stores/loads' of the kernels become plain numpy indexing into the peers' slabs."""

    def __init__(self, d, tmp, rows, cols, channels, rank, world, psf=None, K=0.01):
        self.rank, self.world, self.C = rank, world, channels
        self.H, self.W = rows, cols
        self.Rp, self.Cp = d.next_pow2(rows), d.next_pow2(cols)
        self.Rl, self.Cl = self.Rp // world, self.Cp // world
        self.first_row, self.n_rows = d.row_slab(rank, world, rows)
        self.npairs = (channels + 1) // 2
        self.path = os.path.join(tmp, "slab%d.bin" % rank)
        self.slab = np.memmap(self.path, dtype=np.complex128, mode="w+", shape=(self.npairs, self.Rp, self.Cl))
        self.slab[:] = 0  # rows >= H are never written by anyone; zero BEFORE the handle exchange (a barrier),
        self.slab.flush()  # never inside phase 1, where a faster peer may already be storing its rows here
        self.mm = torch.zeros((channels, 2), dtype=torch.float64)
        if psf is not None:
            self.set_psf(psf, K)

    def set_psf(self, psf, K):
        """Like fdr_shard_set_psf_*: the Wiener slab is built THROUGH the column slab (pair 0 is the scratch of
        the PSF's row spectrum), so a peer that scattered into this slab before the build finished would corrupt it."""
        hp = np.zeros((self.Rp, self.Cp))
        hp[: psf.shape[0], : psf.shape[1]] = psf
        cols = slice(self.rank * self.Cl, (self.rank + 1) * self.Cl)
        slab = np.memmap(self.path, dtype=np.complex128, mode="r+", shape=(self.npairs, self.Rp, self.Cl))
        slab[0] = np.fft.fft(hp, axis=1)[:, cols]
        slab.flush()
        import time
        time.sleep(0.3 * self.rank)  # skew: rank 0 is done long before the last rank
        slab = np.memmap(self.path, dtype=np.complex128, mode="r+", shape=(self.npairs, self.Rp, self.Cl))
        Hs = np.fft.fft(np.array(slab[0]), axis=0)
        self.wf = np.conj(Hs) / (np.abs(Hs) ** 2 + K)
        slab[0] = 0
        slab.flush()

    def export_handle(self):
        return self.path

    def set_peers_from_handles(self, handles):
        self.peers = [np.memmap(h, dtype=np.complex128, mode="r+", shape=(self.npairs, self.Rp, self.Cl)) for h in handles]

    def minmax_tensor(self, device):
        return self.mm

    def phase1(self, rows_u8, stream=0):
        x = rows_u8.astype(np.float64) / 255.0
        for p in range(self.npairs):
            a = x[:, :, 2 * p]
            b = x[:, :, 2 * p + 1] if 2 * p + 1 < self.C else np.zeros_like(a)
            z = np.zeros((self.n_rows, self.Cp), np.complex128)
            z[:, : self.W] = a + 1j * b
            Z = np.fft.fft(z, axis=1)
            for xcol in range(0, self.Cp, self.Cl):            # ROW_OUT_SCATTER
                peer = xcol >> int(np.log2(self.Cl))
                self.peers[peer][p, self.first_row:self.first_row + self.n_rows, :] = Z[:, xcol:xcol + self.Cl]
        for m in self.peers:
            m.flush()

    def phase2(self, stream=0):
        self.slab = np.memmap(self.path, dtype=np.complex128, mode="r+", shape=(self.npairs, self.Rp, self.Cl))
        for p in range(self.npairs):
            Y = np.fft.fft(np.array(self.slab[p]), axis=0) * self.wf
            self.slab[p] = np.fft.ifft(Y, axis=0) * self.Rp
        self.slab.flush()

    def phase3(self, stream=0):
        r0 = self.rank * self.Rl
        self.raw = np.zeros((self.C, self.Rl, self.Cp))
        for p in range(self.npairs):
            row = np.concatenate([np.array(np.memmap(m.filename, dtype=np.complex128, mode="r",
                                                     shape=(self.npairs, self.Rp, self.Cl))[p, r0:r0 + self.Rl, :])
                                  for m in self.peers], axis=1)  # ROW_IN_GATHER
            z = np.fft.ifft(row, axis=1) * self.Cp
            self.raw[2 * p] = z.real
            if 2 * p + 1 < self.C:
                self.raw[2 * p + 1] = z.imag
        self.mm[:, 0] = torch.from_numpy(self.raw.min(axis=(1, 2)))
        self.mm[:, 1] = torch.from_numpy(self.raw.max(axis=(1, 2)))

    def phase4(self, out_rows_u8, stream=0):
        mn, mx = self.mm[:, 0].numpy(), self.mm[:, 1].numpy()
        for c in range(self.C):
            n = (self.raw[c, : self.n_rows, : self.W] - mn[c]) / (mx[c] - mn[c])
            out_rows_u8[:, :, c] = np.clip(np.rint(n * 255.0), 0, 255).astype(np.uint8)


def _worker(rank, world, port, tmp, H, W, C, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    d = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    psf = np.zeros((5, 5))
    psf[2, :] = 0.2
    back = NumpyShard(d, tmp, H, W, C, rank, world)
    drv = d.ShardedRestorer(back)
    drv.set_psf(psf, 0.01)  # fenced: without the cross-rank barrier rank 0's phase 1 lands in rank 1's scratch
    first, n = d.row_slab(rank, world, H)
    out = np.zeros((n, W, C), np.uint8)
    drv.restore_rows(img[first:first + n], out)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, out))
    if rank == 0:
        full = np.zeros((H, W, C), np.uint8)
        for f, o in gathered:
            full[f:f + o.shape[0]] = o
        np.save(result_path, full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W", [(40, 60), (64, 64), (13, 100)])
def test_sharded_driver_world2_gloo(tmp_path, H, W):
    C, world = 3, 2
    port = 29500 + (os.getpid() + H) % 2000
    result = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(world, port, str(tmp_path), H, W, C, result), nprocs=world, join=True)
    got = np.load(result)
    # single-process float64 pipeline (Appendix A of SURVEY.md)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    d = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
    Rp, Cp = d.next_pow2(H), d.next_pow2(W)
    psf = np.zeros((5, 5))
    psf[2, :] = 0.2
    hp = np.zeros((Rp, Cp))
    hp[:5, :5] = psf
    Hs = np.fft.fft2(hp)
    wf = np.conj(Hs) / (np.abs(Hs) ** 2 + 0.01)
    want = np.zeros_like(img)
    for c in range(C):
        g = np.zeros((Rp, Cp))
        g[:H, :W] = img[:, :, c] / 255.0
        f = np.real(np.fft.ifft2(np.fft.fft2(g) * wf)) * Rp * Cp
        n = (f - f.min()) / (f.max() - f.min())
        want[:, :, c] = np.clip(np.rint(n[:H, :W] * 255.0), 0, 255).astype(np.uint8)
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999
