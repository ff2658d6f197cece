"""Timeline of the native pipelined driver on N real GPUs (FDR_SHARD_TIMELINE=1): end times of every step of one restore on
rank 0 and on the slowest rank.  torchrun --nproc-per-node N profiles/shard_timeline.py [link_ctas ...]"""
import json, os, sys
os.environ["FDR_SHARD_TIMELINE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import load_fdr, _load, PKG
fdr = load_fdr()
fd = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
H = W = 16384
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sh = stream.cuda_stream
back = fd.cuda_shard_backend(fdr, H, W, 3, rank, world, lr)
drv = fd.ShardedRestorer(back, device=dev)
drv.set_psf_motion(50, 30.0, 0.01)
d_in = torch.empty((max(back.n_rows, 1), W, 3), dtype=torch.uint8, device=dev)
d_out = torch.zeros_like(d_in)
fdr.synth_rows_device_u8(d_in.data_ptr(), 0xF17E0004, 0, 3, H, W, back.first_row, back.n_rows, sh)
torch.cuda.synchronize()
def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for arg in sys.argv[1:] or ["ce4", "32"]:   # "ceN": copy engines over N streams; "N": link kernel on N CTAs
    if arg.startswith("ce"):
        lc = 0
        back.set_link_ctas(0)
        back.set_ce_streams(int(arg[2:] or 4))
    else:
        lc = int(arg)
        back.set_link_ctas(lc)
    tx = []
    for which in (back.exchange1, back.exchange3):   # the exchange alone, all units
        which(sh)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(3):
            which(sh)
        a1.record(stream)
        torch.cuda.synchronize()
        tx.append(round(maxr(a0.elapsed_time(a1)) / 3, 4))
    for _ in range(5):
        drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh)
    e1.record(stream)
    torch.cuda.synchronize()
    tl = back.timeline()
    if rank == 0:
        print("link", arg, "exchange1/3 alone ms", tx, "ms/step %.3f" % (e0.elapsed_time(e1) / 10), json.dumps(tl), flush=True)
    dist.barrier()
back.close()
dist.destroy_process_group()
