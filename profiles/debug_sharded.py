"""Multi-GPU diagnosis of the row-sharded path: every rank restores the whole synthetic image on its own GPU
(unsharded plan) and compares ITS rows of the sharded result; the sharded restore runs non-pipelined twice and
pipelined once (determinism).  torchrun --nproc-per-node N profiles/debug_sharded.py [H]"""
import os, sys
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
seed = 0xF17E0004
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sh = stream.cuda_stream
back = fd.cuda_shard_backend(fdr, H, W, 3, rank, world, lr)
drv = fd.ShardedRestorer(back, device=dev)
drv.set_psf_motion(50, 30.0, 0.01)
n_rows, first = back.n_rows, back.first_row
d_in = torch.empty((n_rows, W, 3), dtype=torch.uint8, device=dev)
fdr.synth_rows_device_u8(d_in.data_ptr(), seed, 0, 3, H, W, first, n_rows, sh)
outs = {}
if os.environ.get("DEBUG_BENCH_LIKE"):
    d_tmp = torch.zeros_like(d_in)
    for _ in range(13):  # warm-up + timed steps of bench.py: back to back, no host synchronisation in between
        drv.restore_rows(d_in.data_ptr(), d_tmp.data_ptr(), sh)
    dist.barrier()
    torch.cuda.synchronize()
    outs["after13pipe"] = d_tmp.clone()
for name, pipe in (("plain1", False), ("plain2", False), ("pipe1", True), ("pipe2", True), ("plain3", False)):
    d_out = torch.zeros_like(d_in)
    drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh, pipeline_pairs=pipe)
    torch.cuda.synchronize()
    dist.barrier()
    outs[name] = d_out
whole = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
ref = torch.empty_like(whole)
fdr.synth_images_device_u8(whole.data_ptr(), seed, 0, 1, 3, H, W, sh)
with fdr.Plan(H, W, 3, 1, lr) as plan:
    plan.set_psf_motion(50, 30.0, 0.01)
    plan.restore_images_device_u8(whole.data_ptr(), ref.data_ptr(), 1, sh)
    torch.cuda.synchronize()
mine = ref[first:first + n_rows]
msg = []
for name, o in outs.items():
    d = (o.to(torch.int16) - mine.to(torch.int16)).abs()
    msg.append("%s: nz %d max %d" % (name, int((d > 0).sum()), int(d.max())))
same = bool((outs["plain1"] == outs["plain2"]).all())
print("rank %d blocks=%s | %s | plain1==plain2 %s" % (rank, os.environ.get("FDR_COL_BLOCKS", "default"), " | ".join(msg), same), flush=True)
dist.barrier()
dist.destroy_process_group()
