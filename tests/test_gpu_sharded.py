"""GPU tests of the row-sharded path (include/fdr_b200.h fdr_shard_*): `world` shards live in ONE
process on ONE device (peer pointers are plain device pointers), the phases of all ranks run in
order with a device synchronise where the real run has a cross-rank barrier.  The result must
equal the unsharded plan and pass the oracle gates."""
import numpy as np
import pytest

from conftest import PKG, _load, u8_gate

pytestmark = pytest.mark.gpu
K = 0.01


def run_emulated(gpu, torch, images_hwc, world, psf_len, psf_ang):
    H, W, C = images_hwc.shape
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(images_hwc).to(dev)
    d_out = torch.zeros_like(d_in)
    dist_mod = _load("fdr_dist", PKG + "/fdr_dist.py")
    shards = [gpu.Shard(H, W, C, g, world, 0) for g in range(world)]
    try:
        slabs = [s.local_slab()[0] for s in shards]
        for g, s in enumerate(shards):
            assert (s.first_row, s.n_rows) == dist_mod.row_slab(g, world, H)
            s.set_peers(slabs)
            s.set_psf_motion(psf_len, psf_ang, K)
        stream = torch.cuda.Stream(device=dev)
        sh = stream.cuda_stream
        rowbytes = W * C
        for s in shards:
            s.phase1(d_in.data_ptr() + s.first_row * rowbytes, sh)
        torch.cuda.synchronize()
        for s in shards:
            s.phase2(sh)
        torch.cuda.synchronize()
        for s in shards:
            s.phase3(sh)
        torch.cuda.synchronize()
        mms = [dist_mod.device_tensor(s.minmax_ptr(), (C, 2), dev) for s in shards]
        allmm = torch.stack(mms)
        gmin, gmax = allmm[:, :, 0].min(0).values, allmm[:, :, 1].max(0).values
        for t in mms:
            t[:, 0] = gmin
            t[:, 1] = gmax
        torch.cuda.synchronize()
        for s in shards:
            s.phase4(d_out.data_ptr() + s.first_row * rowbytes, sh)
        torch.cuda.synchronize()
        launches = sum(s.last_launch_count() for s in shards)
        return d_out.cpu().numpy(), launches
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("H,W,world", [(200, 320, 2), (200, 320, 4), (256, 512, 8), (64, 1024, 2), (1000, 40, 4), (33, 70, 8),
                                       (8192, 64, 2), (16384, 128, 8)])
def test_sharded_equals_single_gpu_and_oracle(gpu, oracle, H, W, world):
    torch = pytest.importorskip("torch")
    img = np.ascontiguousarray(np.transpose(oracle.synth_image_u8(5, 0, H, W), (1, 2, 0)))
    got, launches = run_emulated(gpu, torch, img, world, 9, 30.0)
    assert launches > 0
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(9, 30.0, K)
        single = p.restore_images_u8(img[None])[0]
    assert np.array_equal(got, single), u8_gate(got, single)
    planes = [img[:, :, c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
    want, _ = oracle.restore_image_u8(planes, oracle.port().motion_psf(9, 30.0), K)
    exact, off1, worse = u8_gate(got, want)
    assert worse == 0 and (exact + off1) / got.size >= 0.999


@pytest.mark.parametrize("H,W,world", [(192, 256, 2), (8192, 64, 2), (16384, 4096, 2)])
def test_sharded_per_pair_phases(gpu, oracle, H, W, world):
    """fdr_shard_phase*_pairs: running the two plane pairs separately, on two concurrent streams (as
    the pipelined driver does), gives the same bytes as the all-pairs phases -- also for the long-column
    schemes (8192 / 16384 rows) at a realistic slab width."""
    torch = pytest.importorskip("torch")
    img = np.ascontiguousarray(np.transpose(oracle.synth_image_u8(5, 1, H, W), (1, 2, 0)))
    want, _ = run_emulated(gpu, torch, img, world, 9, 30.0)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(img).to(dev)
    d_out = torch.zeros_like(d_in)
    dist_mod = _load("fdr_dist", PKG + "/fdr_dist.py")
    shards = [gpu.Shard(H, W, 3, g, world, 0) for g in range(world)]
    try:
        slabs = [s.local_slab()[0] for s in shards]
        for s in shards:
            assert s.npairs == 2
            s.set_peers(slabs)
            s.set_psf_motion(9, 30.0, K)
        st = torch.cuda.Stream(device=dev)
        sts = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        rb = W * 3
        for ph in (1, 2, 3):
            for pair in (1, 0):  # order between pairs must not matter; the two pairs run concurrently
                for s in shards:
                    if ph == 1:
                        s.phase1(d_in.data_ptr() + s.first_row * rb, sts[pair].cuda_stream, pair=pair)
                    elif ph == 2:
                        s.phase2(sts[pair].cuda_stream, pair=pair)
                    else:
                        s.phase3(sts[pair].cuda_stream, pair=pair)
            torch.cuda.synchronize()
        mms = [dist_mod.device_tensor(s.minmax_ptr(), (3, 2), dev) for s in shards]
        allmm = torch.stack(mms)
        gmin, gmax = allmm[:, :, 0].min(0).values, allmm[:, :, 1].max(0).values
        for t in mms:
            t[:, 0] = gmin
            t[:, 1] = gmax
        torch.cuda.synchronize()
        for s in shards:
            s.phase4(d_out.data_ptr() + s.first_row * rb, st.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)
    finally:
        for s in shards:
            s.close()


def test_sharded_4096_world8(gpu, oracle):
    """A BASELINE-sized plane through the sharded kernels (column slabs of 512 columns)."""
    torch = pytest.importorskip("torch")
    H = W = 4096
    img = np.ascontiguousarray(np.transpose(oracle.synth_image_u8(2, 0, H, W), (1, 2, 0)))
    got, _ = run_emulated(gpu, torch, img, 8, 50, 30.0)
    with gpu.Plan(H, W, 3) as p:
        p.set_psf_motion(50, 30.0, K)
        single = p.restore_images_u8(img[None])[0]
    assert np.array_equal(got, single)


def test_synth_rows_match_whole_image(gpu, oracle):
    torch = pytest.importorskip("torch")
    H, W = 96, 130
    whole = np.transpose(oracle.synth_image_u8(4, 3, H, W), (1, 2, 0))
    d = torch.zeros((40, W, 3), dtype=torch.uint8, device="cuda")
    gpu.synth_rows_device_u8(d.data_ptr(), 0xF17E0000 + 4, 3, 3, H, W, 17, 40, 0)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), whole[17:57])
