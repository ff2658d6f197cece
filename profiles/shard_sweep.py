"""Row-sharded 16384^2 (BASELINE configs[4]) on N real GPUs: variants of the driver in ONE process group, so that one box
acquisition answers several questions.  torchrun --nproc-per-node N profiles/shard_sweep.py [H] [out.json]
Per variant: ms per image (CUDA events, max over ranks), plus the per-phase times of the serial schedule, the cost of the two
barrier flavours, and every rank's rows against an unsharded plan on its own GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import load_fdr, _load, PKG
fdr = load_fdr()
fd = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
out_path = sys.argv[2] if len(sys.argv) > 2 else None
STEPS = int(os.environ.get("SWEEP_STEPS", "10"))
seed = 0xF17E0004
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sh = stream.cuda_stream
res = {"world": world, "H": H, "W": W, "steps": STEPS, "variants": {}}


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_mode(half):
    os.environ["FDR_SHARD_HALF"] = "1" if half else "0"
    back = fd.cuda_shard_backend(fdr, H, W, 3, rank, world, lr)
    drv = fd.ShardedRestorer(back, device=dev)
    drv.set_psf_motion(50, 30.0, 0.01)
    n_rows, first = back.n_rows, back.first_row
    d_in = torch.empty((max(n_rows, 1), W, 3), dtype=torch.uint8, device=dev)
    d_out = torch.zeros_like(d_in)
    fdr.synth_rows_device_u8(d_in.data_ptr(), seed, 0, 3, H, W, first, n_rows, sh)
    torch.cuda.synchronize()
    tag = ("staged" if back.staged else "half") if half else "pair"

    def timed(name, peer_sync, pipe, ctas):
        drv.peer_sync = peer_sync
        back.set_row_ctas(ctas)
        for _ in range(3):
            drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh, pipeline_pairs=pipe)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(STEPS):
            drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh, pipeline_pairs=pipe)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = maxr(e0.elapsed_time(e1)) / STEPS
        res["variants"]["%s.%s" % (tag, name)] = ms
        if rank == 0:
            print("%s.%s: %.3f ms" % (tag, name, ms), flush=True)

    drv.native = False
    if os.environ.get("SWEEP_FULL"):
        timed("serial.nccl", False, False, 0)
        timed("pipe.nccl", False, True, 0)
        for ctas in ((0,) if back.staged else (0, 74, 120)):
            timed("pipe.peer.ctas%d" % ctas, True, True, ctas)
    timed("serial.peer", True, False, 0)
    back.set_row_ctas(0)
    drv.peer_sync = True
    drv.native = True
    for lc in ((0, 8, 16, 24, 32) if back.staged else (32,)):
        back.set_link_ctas(lc)
        timed("native.link%d" % lc, True, True, 0)
    back.set_link_ctas(16)
    # per-phase times of the serial schedule
    ph = np.zeros(7)
    reps = 5
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        ev[0].record(stream); back.phase1(d_in.data_ptr(), sh); x0 = torch.cuda.Event(enable_timing=True); x0.record(stream)
        back.exchange1(sh); ev[1].record(stream)
        drv.barrier(set_index=13, stream=sh)
        ev[2].record(stream); back.phase2(sh); x1 = torch.cuda.Event(enable_timing=True); x1.record(stream)
        back.exchange3(sh); ev[3].record(stream)
        drv.barrier(set_index=14, stream=sh)
        ev[4].record(stream); back.phase3(sh); ev[5].record(stream)
        drv._reduce_minmax()
        ev[6].record(stream); back.phase4(d_out.data_ptr(), sh); ev[7].record(stream)
        torch.cuda.synchronize()
        ph += np.array([ev[0].elapsed_time(x0), x0.elapsed_time(ev[1]), ev[2].elapsed_time(x1), x1.elapsed_time(ev[3]),
                        ev[4].elapsed_time(ev[5]), ev[6].elapsed_time(ev[7]), ev[0].elapsed_time(ev[7])])
    ph = [maxr(float(x) / reps) for x in ph]
    res["variants"]["%s.phases_ms" % tag] = dict(zip(["phase1", "exchange1", "phase2", "exchange3", "phase3", "phase4", "total_serial"], ph))
    if rank == 0:
        print(tag, "phases", ph, flush=True)
    if back.staged:   # the link kernel alone: time per exchange (all units) against the number of CTAs
        for lc in (0, 8, 16, 32):
            back.set_link_ctas(lc)
            tx = []
            for which in (back.exchange1, back.exchange3):
                which(sh)
                torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3):
                    which(sh)
                e1.record(stream)
                torch.cuda.synchronize()
                tx.append(maxr(e0.elapsed_time(e1)) / 3)
            res["variants"]["%s.exchange_ms.link%d" % (tag, lc)] = tx
            if rank == 0:
                print(tag, "exchange1/3 ms at", lc, "CTAs:", tx, flush=True)
        back.set_link_ctas(16)
    # barrier cost
    for name, ps in (("peer", True), ("nccl", False)):
        drv.peer_sync = ps
        for _ in range(5):
            drv.barrier(set_index=12, stream=sh)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            drv.barrier(set_index=12, stream=sh)
        e1.record(stream)
        torch.cuda.synchronize()
        res["variants"]["%s.barrier_us.%s" % (tag, name)] = maxr(e0.elapsed_time(e1)) / 50 * 1e3
    drv.peer_sync = True
    assert not back.sync_timed_out(sh)
    # parity of this rank's rows against the unsharded plan on this GPU (same arithmetic family: half-plane on both, or neither)
    drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh)
    torch.cuda.synchronize()
    os.environ["FDR_HALF"] = "1" if half else "0"
    whole = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    ref = torch.empty_like(whole)
    fdr.synth_images_device_u8(whole.data_ptr(), seed, 0, 1, 3, H, W, sh)
    with fdr.Plan(H, W, 3, 1, lr) as plan:
        plan.set_psf_motion(50, 30.0, 0.01)
        plan.restore_images_device_u8(whole.data_ptr(), ref.data_ptr(), 1, sh)
        torch.cuda.synchronize()
    d = (d_out[:n_rows].to(torch.int16) - ref[first:first + n_rows].to(torch.int16)).abs()
    cnt = torch.tensor([int((d == 1).sum()), int((d > 1).sum())], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt)
    res["variants"]["%s.vs_unsharded" % tag] = {"off_by_1": int(cnt[0].item()), "off_by_more": int(cnt[1].item()), "pixels": H * W * 3}
    del whole, ref, d
    dist.barrier()
    back.close()


for half, staged in ((True, True), (True, False)) + (((False, False),) if os.environ.get("SWEEP_PAIR") else ()):
    os.environ["FDR_SHARD_STAGED"] = "1" if staged else "0"
    run_mode(half)
if rank == 0:
    print(json.dumps(res))
    if out_path:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)
dist.barrier()
dist.destroy_process_group()
