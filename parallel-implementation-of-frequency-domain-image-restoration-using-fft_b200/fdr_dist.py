"""Partitioner for one 8 x B200 box (torch.distributed plumbing over NCCL/NVLink; every kernel is
in lib/libfdr_b200.so).

Two granularities:
  * batches / colour planes -- independent: images_for_rank() gives each rank a contiguous block
    of images; no data-path collective (bench.py --gpus N).
  * one very large image -- ShardedRestorer: row slabs per rank, the transposes fused into the row
    passes as peer stores/loads, cross-rank barriers and the 2-float min/max all-reduce through
    torch.distributed.  Replaces the reference's MPI driver loop (mpi.cpp:95-111 +
    fft_mpi.cpp:311-470).
"""
import os

import torch
import torch.distributed as dist


def next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


def row_slab(rank, world, image_rows):
    """Image rows [first, first+n) owned by `rank`: the padded rows are split evenly
    (fft_mpi.cpp:89-100 calculate_distribution on a power of two has no remainder) and the part
    beyond the image is zero padding nobody stores.  Must equal fdr_shard_geometry()."""
    rl = next_pow2(image_rows) // world
    first = rank * rl
    last = min(first + rl, image_rows)
    return first, max(0, last - first)


def images_for_rank(rank, world, n_images):
    """Contiguous block of images for `rank` (first `n_images % world` ranks get one more)."""
    base, rem = divmod(n_images, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class _DevView:
    """__cuda_array_interface__ view over raw device memory, for torch.as_tensor (zero copy)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape, device, typestr="<f4"):
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class ShardedRestorer:
    """Per-rank driver of the row-sharded restoration.  `backend` is an fdr.Shard (CUDA) or any
    object with the same phase API (the CPU gloo test passes a numpy stand-in).

    Pipeline units: a backend splits the image into `npairs` independent units (plane pairs, or single
    planes in half-plane mode).  With more than one unit and a CUDA backend the units run on their own
    streams -- phase 1 | barrier | phase 2 | barrier | phase 3 per unit -- so that unit u's NVLink-bound
    row phases overlap unit u-1's HBM-bound column phase; every rank issues the collectives in the same
    order.  The reference runs its channels strictly one after the other (mpi.cpp:95-111)."""

    def __init__(self, backend, group=None, device=None):
        self.b = backend
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert self.rank == backend.rank and self.world == backend.world
        # exchange the slab handles (64-byte CUDA IPC handles, or whatever the backend exports)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, backend.export_handle(), group=group)
        else:
            handles[0] = backend.export_handle()
        backend.set_peers_from_handles(handles)
        self._mm = backend.minmax_tensor(device)      # [channels][2] view, all-reduced in place
        self._flag = torch.zeros(1, dtype=torch.float32, device=self._mm.device)
        self._side = None                             # side streams + flags of the unit pipeline
        self._cuda = self._mm.is_cuda
        # per-image synchronisation over peer memory instead of NCCL collectives (FDR_SHARD_PEER_SYNC=0 selects NCCL)
        self.peer_sync = (self._cuda and self.world > 1 and hasattr(backend, "peer_barrier")
                          and os.environ.get("FDR_SHARD_PEER_SYNC", "1") != "0")
        self.native = os.environ.get("FDR_SHARD_NATIVE", "1") != "0"   # pipelined driver inside the library

    def set_psf_motion(self, length, angle_deg, K):
        """PSF + Wiener factor on every rank, then a cross-rank fence: the build uses the rank's column slab as
        scratch, so no peer may start scattering into it (phase 1 of a restore) before every rank is done."""
        self.b.set_psf_motion(length, angle_deg, K)
        self._fence()

    def set_psf(self, psf, K):
        self.b.set_psf(psf, K)
        self._fence()

    def _fence(self):
        if self.world > 1:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def close(self):
        """Collective tear-down: every rank unmaps its peers' slabs, THEN (after a cross-rank fence) frees its own -- CUDA IPC
        leaves freeing exported memory that a peer still has mapped undefined (the reference's MPI_Finalize ordering,
        mpi.cpp:124-127, has no such constraint because it never shares device memory)."""
        if hasattr(self.b, "close_peers"):
            self.b.close_peers()
        self._fence()
        self._mm = None
        self.b.close()

    def barrier(self, flag=None, set_index=0, stream=None):
        """Stream-ordered cross-rank barrier: flags in peer memory (fdr_shard_barrier) when the backend has them and
        `peer_sync` is on, else a 1-element all-reduce on the current stream through torch.distributed."""
        if self.world <= 1:
            return
        if self.peer_sync:
            self.b.peer_barrier(set_index, torch.cuda.current_stream(self._mm.device).cuda_stream if stream is None else stream)
        else:
            dist.all_reduce(self._flag if flag is None else flag, group=self.group)

    def _reduce_minmax(self):
        """Global extrema of every padded plane (also the barrier that frees the slabs for the next image)."""
        if self.world <= 1:
            return
        if self.peer_sync:
            self.b.minmax_allreduce(torch.cuda.current_stream(self._mm.device).cuda_stream)
            return
        if getattr(self.b, "minmax_negated", False):
            dist.all_reduce(self._mm, op=dist.ReduceOp.MIN, group=self.group)  # (min, -max): one collective
            return
        mn = self._mm[:, 0].contiguous()
        mx = self._mm[:, 1].contiguous()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        self._mm[:, 0].copy_(mn)
        self._mm[:, 1].copy_(mx)

    def _exchange(self, which, stream, unit=None):
        """Staged backends move the column blocks with their own link kernels (fdr_shard_exchange1/3); fused ones did it
        inside the row passes."""
        f = getattr(self.b, "exchange1" if which == 1 else "exchange3", None)
        if f is not None:
            f(stream, pair=unit)

    def _restore_rows_unit_pipeline(self, d_in_rows, main):
        """Every unit on its own side stream, phases interleaved across units in issue order."""
        b = self.b
        n = b.npairs
        if self._side is None or len(self._side) != n:
            dev = self._mm.device
            self._side = [(torch.cuda.Stream(device=dev), torch.zeros(1, dtype=torch.float32, device=dev)) for _ in range(n)]
        for st, _ in self._side:
            st.wait_stream(main)
        for phase in (1, 2, 3):
            for unit, (st, flag) in enumerate(self._side):
                with torch.cuda.stream(st):
                    if phase == 1:
                        b.phase1(d_in_rows, st.cuda_stream, pair=unit)
                        self._exchange(1, st.cuda_stream, unit)
                    elif phase == 2:
                        b.phase2(st.cuda_stream, pair=unit)
                        self._exchange(3, st.cuda_stream, unit)
                    else:
                        b.phase3(st.cuda_stream, pair=unit)
                    if phase < 3:
                        self.barrier(flag, set_index=2 * unit + phase - 1, stream=st.cuda_stream)
        for st, _ in self._side:
            main.wait_stream(st)

    def restore_rows(self, d_in_rows, d_out_rows, stream=None, pipeline_pairs=True):
        """stream: raw CUDA stream handle of torch's CURRENT stream (default: looked up).  The barriers and the min/max
        all-reduce are torch.distributed calls, which order themselves against the current stream only, so the phases
        must run on that same stream; any other handle (including 0, the legacy stream) is rejected."""
        b = self.b
        if self._cuda:
            cur = torch.cuda.current_stream(self._mm.device)
            if stream is None:
                stream = cur.cuda_stream
            elif stream != cur.cuda_stream:
                raise ValueError("restore_rows: `stream` must be torch's current stream (set it with torch.cuda.stream / "
                                 "set_stream); the cross-rank barriers are ordered against that stream only")
        elif stream is None:
            stream = 0
        if pipeline_pairs and self.peer_sync and self.native and b.npairs <= 6 and hasattr(b, "restore_rows_native"):
            b.restore_rows_native(d_in_rows, d_out_rows, stream)   # the whole pipeline in one native call (fdr_shard_restore_rows)
            return
        if (pipeline_pairs and self._cuda and self.world > 1 and getattr(b, "supports_pair_pipeline", False) and 2 <= b.npairs <= 6):   # flag sets 0..11 belong to the units
            self._restore_rows_unit_pipeline(d_in_rows, cur)
        else:
            b.phase1(d_in_rows, stream)
            self._exchange(1, stream)
            self.barrier(set_index=13, stream=stream)   # every slab has received all its columns
            b.phase2(stream)
            self._exchange(3, stream)
            self.barrier(set_index=14, stream=stream)   # every slab / staging plane holds the filtered, column-inverted data
            b.phase3(stream)
        self._reduce_minmax()
        b.phase4(d_out_rows, stream)


def cuda_shard_backend(fdr, rows, cols, channels, rank, world, device_index):
    """fdr.Shard plus the two hooks ShardedRestorer needs."""
    sh = fdr.Shard(rows, cols, channels, rank, world, device_index)
    sh.minmax_tensor = lambda device: device_tensor(sh.minmax_ptr(), (channels, 2), device or torch.device("cuda", device_index))
    sh.supports_pair_pipeline = True
    sh.set_minmax_negated(True)   # one all-reduce(MIN) over (min, -max) instead of two collectives + four copies
    return sh
