#!/usr/bin/env python
"""bench.py -- Mpixel/s deblurred (FFT -> Wiener -> IFFT -> normalise -> 8-bit pack).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

One "step" restores one batch of synthetic 8-bit BGR images (default workload = BASELINE.json
configs[3]: 256 x 2048x2048x3, counter-hash pixels, psf 50/30, K = 0.01).  Inputs are resident in
HBM before the timed region; `value` is whole-job Mpixel/s (max time over ranks); `e2e` is the
same metric through the host-buffer C-ABI call (pinned host -> device -> pinned host inside the
timed region).  Rank 0 prints ONE JSON line.  Multi-GPU: independent image batches per rank, no
data-path collective (weak scaling).  torch is used only for device memory, streams, events and
torch.distributed plumbing; every kernel is ours (lib/libfdr_b200.so via ctypes).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
# the sharded driver keeps compute, link and barrier streams apart: give every stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
PKG = os.path.join(ROOT, "parallel-implementation-of-frequency-domain-image-restoration-using-fft_b200")

WORKLOADS = {
    # name: (config_index, images, H, W, psf_len, psf_angle)
    "batch256x2048": (3, 256, 2048, 2048, 50, 30.0),
    "rgb4096": (2, 1, 4096, 4096, 50, 30.0),
    "rgb16384": (4, 1, 16384, 16384, 50, 30.0),
    "car": (1, 1, 330, 640, 40, 45.0),
    "cat": (0, 1, 782, 1920, 50, 30.0),
}
K_WIENER = 0.01
CHILD_ENTRY = os.path.abspath(__file__)   # what sharded_leg_isolated launches (the CPU test harness points it at itself)
CONTRACT_BYTES_PER_CHANNEL_PIXEL = 53.0


_REAL_STDOUT = None


def claim_stdout():
    """The reference's translation units (oracle/_ref) print their profile blocks on stdout from C++; the contract is ONE JSON
    line on stdout.  Keep the real stdout aside for that line and send everything else (C and Python) to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def _trace(msg):
    """FDR_BENCH_TRACE=1: stage markers on stderr (debugging aid)."""
    if os.environ.get("FDR_BENCH_TRACE"):
        print("[bench %.1fs] %s" % (time.perf_counter(), msg), file=sys.stderr, flush=True)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_restore_images(O, mode, cfg_idx, first, count, H, W, psf, threads):
    """Reference CPU path on `count` synthetic images: serial.cpp:33-39 loop per image."""
    ref = O.ref() if O.have_ref() else None
    import numpy as np
    if ref is not None and mode == "openmp":
        ref.set_threads(threads)
    t_total = 0.0
    for i in range(first, first + count):
        img = O.synth_image_u8(cfg_idx, i, H, W)
        planes = [O.pad_pow2(img[c].astype(np.float32) * np.float32(1.0 / 255.0)) for c in range(3)]
        t0 = time.perf_counter()
        for pl in planes:
            if ref is not None:
                ref.wiener(pl, psf, K_WIENER, mode)
            else:
                O.port().wiener_deblur(pl, psf, K_WIENER)
        t_total += time.perf_counter() - t0
    return t_total


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, compiled unmodified
    from the reference sources; openmp mode with every host thread), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O = _load("fdr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    cfg_idx, images, H, W, plen, pang = WORKLOADS[args.workload]
    psf = O.port().motion_psf(plen, pang)
    have_ref = O.have_ref()
    threads = os.cpu_count() or 1
    mode = args.ref_mode if have_ref else "serial"
    sample = max(1, min(images, args.ref_images))
    if mode == "mpi":
        # the reference's MPI backend over the single-node stand-in (oracle/mpi_standin), ranks = host cores (max 16)
        import numpy as np
        ranks = max(1, min(threads, 16))
        t = 0.0
        for s in range(args.warmup + args.steps):
            tt = 0.0
            for i in range(s * sample, (s + 1) * sample):
                img = O.synth_image_u8(cfg_idx, i, H, W)
                planes = np.stack([O.pad_pow2(img[c].astype(np.float32) * np.float32(1.0 / 255.0)) for c in range(3)])
                tt += O.ref_mpi_wiener(planes, psf, K_WIENER, ranks)[1] * 1e-3
            if s >= args.warmup:
                t += tt
        threads = ranks
    else:
        for _ in range(args.warmup):
            cpu_restore_images(O, mode, cfg_idx, 0, 1, H, W, psf, threads)
        t = 0.0
        for s in range(args.steps):
            t += cpu_restore_images(O, mode, cfg_idx, s * sample, sample, H, W, psf, threads)
    mpx = sample * H * W * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": "Mpixel/s deblurred (FFT->Wiener->IFFT->normalise->8-bit pack)",
        "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_step_sample": sample, "image": [H, W, 3], "psf": [plen, pang],
                   "K": K_WIENER},
        "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": threads if mode in ("openmp", "mpi") else 1,
                         "kind": "reference" if have_ref else "port",
                         "sample": "%d image(s) of %dx%dx3 per step, reference %s mode (fft_%s.cpp compiled unmodified)"
                                   % (sample, H, W, mode, mode)},
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def pcie_roofline(torch, dist, dev, world, mb=256, reps=4):
    """Measured concurrent pinned H2D + D2H rate of this rank's GPU while every rank does the same (the end-to-end bound)."""
    n = mb << 20
    hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    for _ in range(1):
        with torch.cuda.stream(s1):
            a.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            a.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    each = n * reps / dt / 1e9
    del hin, hout, a, b
    return {"GBps_each_way_per_gpu": each, "GBps_each_way_all_gpus": each * world,
            "how": "%d MiB pinned H2D and D2H concurrently on two streams, x%d, all %d ranks at once, slowest rank" % (mb, reps, world)}


def sharded_parity(fdr, torch, got, mode, cfg_idx, H, W, plen, pang, seed, dev, local_rank, sh):
    """Rank 0: the gathered result against (mode "oracle") plane 0 of the WHOLE image restored by the reference's CPU code at
    full size -- oracle/_ref openmp mode, which agrees with its serial mode to 1e-7; the serial port otherwise
    (fft_serial.cpp:141-261) -- or (mode "self") the whole image restored by this library's single-GPU path."""
    import numpy as np
    if mode == "oracle":
        O = _load("fdr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
        psf = O.port().motion_psf(plen, pang)
        img0 = O.synth_image_u8(cfg_idx, 0, H, W, channels=1)[0]
        pl = O.pad_pow2(img0.astype(np.float32) * np.float32(1.0 / 255.0))
        t0 = time.perf_counter()
        if O.have_ref():
            O.ref().set_threads(os.cpu_count() or 1)
            norm = O.ref().wiener(pl, psf, K_WIENER, "openmp")[:H, :W]
            how = "reference fft_openmp.cpp compiled unmodified (oracle/_ref), %d threads" % (os.cpu_count() or 1)
        else:
            norm = O.port().wiener_deblur(pl, psf, K_WIENER)["norm"][:H, :W]
            how = "oracle port (serial)"
        t_cpu = time.perf_counter() - t0
        want = O.port().pack_u8(norm)
        against = "%s on the whole %dx%d plane 0, %.1f s" % (how, H, W, t_cpu)
    else:
        whole = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        ref_out = torch.empty_like(whole)
        fdr.synth_images_device_u8(whole.data_ptr(), seed, 0, 1, 3, H, W, sh)
        with fdr.Plan(H, W, 3, 1, local_rank) as plan:
            plan.set_psf_motion(plen, pang, K_WIENER)
            plan.restore_images_device_u8(whole.data_ptr(), ref_out.data_ptr(), 1, sh)
            torch.cuda.synchronize()
        want = ref_out.cpu().numpy()
        against = "single-GPU path of this library, whole image"
        del whole, ref_out
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return {"against": against, "pixels": int(d.size), "exact": int((d == 0).sum()), "off_by_1": int((d == 1).sum()),
            "off_by_more": int((d > 1).sum()), "frac_within_1": float((d <= 1).mean())}


def sharded_measure(args, fdr, torch, dist, world, rank, local_rank, dev, cfg_idx, H, W, plen, pang, seed, steps, warmup,
                    want_e2e=True, parity_mode="oracle"):
    """One image row-sharded over `world` GPUs (BASELINE configs[4]); strong scaling.  The transposes of the reference's
    MPI_Alltoallv (fft_mpi.cpp:170-279, 284-307) are peer stores/loads fused into the row passes; barriers and the min/max
    all-reduce run over flags in peer memory; torch.distributed (NCCL) only exchanges the IPC handles and fences the PSF
    build.  Collective: every rank calls it; returns the result dict on every rank (parity only on rank 0)."""
    import numpy as np
    fd = _load("fdr_dist", os.path.join(PKG, "fdr_dist.py"))
    stream = torch.cuda.current_stream(dev)
    sh = stream.cuda_stream
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~1 s to start: begin before set-up and warm-up so the timed region is covered
    back = fd.cuda_shard_backend(fdr, H, W, 3, rank, world, local_rank)
    drv = fd.ShardedRestorer(back, device=dev)
    drv.set_psf_motion(plen, pang, K_WIENER)  # builds the Wiener slab on every rank, then fences across ranks
    n_rows, first = back.n_rows, back.first_row
    d_in = torch.empty((max(n_rows, 1), W, 3), dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(d_in)
    fdr.synth_rows_device_u8(d_in.data_ptr(), seed, 0, 3, H, W, first, n_rows, sh)
    torch.cuda.synchronize()

    def step():
        drv.restore_rows(d_in.data_ptr(), d_out.data_ptr(), sh)

    # Every restore contains cross-rank barriers, so every rank must run the SAME number of them: loop counts below are derived
    # from a duration all-reduced over the ranks, never from a rank's own clock.
    t_w0 = time.perf_counter()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    tw = torch.tensor([(time.perf_counter() - t_w0) / max(1, warmup)], dtype=torch.float64, device=dev)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    per_step_s = max(float(tw.item()), 1e-4)
    n_fill = int(min(3000, max(0, 1.5 / per_step_s)))   # ~1.5 s of the same load: the clock sampler needs ~1 s to start
    for i in range(n_fill):
        step()
        if i % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    # a longer busy stretch of the same step so that the 200 ms sampler sees the clocks under this load
    n_busy = int(min(3000, max(20, 1.0 / max(ms_step * 1e-3, 1e-4))))   # identical on every rank (ms_step is all-reduced)
    for i in range(n_busy):
        step()
        if i % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dist.barrier()

    # per-phase device times of the serial schedule (separate pass; the timed region above runs the pipelined driver)
    ph = np.zeros(7)
    reps = 3
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        ev[0].record(stream)
        back.phase1(d_in.data_ptr(), sh)
        ev[1].record(stream)
        back.exchange1(sh)
        ev[2].record(stream)
        drv.barrier(set_index=13, stream=sh)
        ev[3].record(stream)
        back.phase2(sh)
        ev[4].record(stream)
        back.exchange3(sh)
        ev[5].record(stream)
        drv.barrier(set_index=14, stream=sh)
        ev[6].record(stream)
        back.phase3(sh)
        ev[7].record(stream)
        drv._reduce_minmax()
        ev[8].record(stream)
        back.phase4(d_out.data_ptr(), sh)
        ev[9].record(stream)
        torch.cuda.synchronize()
        ph += np.array([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[3].elapsed_time(ev[4]), ev[4].elapsed_time(ev[5]),
                        ev[6].elapsed_time(ev[7]), ev[8].elapsed_time(ev[9]), ev[0].elapsed_time(ev[9])])
    pt = torch.tensor(ph / reps, dtype=torch.float64, device=dev)
    dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ph = [float(x) for x in pt.tolist()]

    # end to end: pinned host rows -> device -> restore -> pinned host rows every step, double-buffered so that the copies of
    # neighbouring steps overlap the restore (three streams)
    e2e = None
    if want_e2e:
        hin = torch.empty(d_in.shape, dtype=torch.uint8, pin_memory=True)
        hout = [torch.empty(d_in.shape, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        hin.copy_(d_in)
        dins = [d_in, torch.empty_like(d_in)]
        douts = [d_out, torch.empty_like(d_out)]
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        n_e = max(4, args.e2e_steps * 4)

        def e2e_loop(n):
            for i in range(n):
                b = i & 1
                with torch.cuda.stream(s_in):
                    if i >= 2:
                        s_in.wait_event(ev_cmp[b])       # the restore that read dins[b] two steps ago
                    dins[b].copy_(hin, non_blocking=True)
                    ev_in[b].record(s_in)
                stream.wait_event(ev_in[b])
                if i >= 2:
                    stream.wait_event(ev_out[b])         # douts[b] has been copied out
                drv.restore_rows(dins[b].data_ptr(), douts[b].data_ptr(), sh)
                ev_cmp[b].record(stream)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b])
                    hout[b].copy_(douts[b], non_blocking=True)
                    ev_out[b].record(s_out)
            torch.cuda.synchronize()

        e2e_loop(2)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_loop(n_e)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": H * W * n_e / float(tt.item()) / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": int(hin.numel()) * world, "d2h_bytes_per_step": int(hout[0].numel()) * world,
               "steps": n_e, "ms_per_step": float(tt.item()) / n_e * 1e3,
               "api": "ShardedRestorer.restore_rows on pinned host row slabs, H2D / restore / D2H of neighbouring steps overlapped (double-buffered)"}
        e2e["matches_device"] = bool(torch.equal(hout[(n_e - 1) & 1][:n_rows], d_out[:n_rows].cpu()) if n_rows else True)
        del hin, hout, dins, douts

    # parity: plane 0 (B) of the WHOLE image against the reference's CPU code at full size (oracle/_ref openmp mode, which
    # agrees with its serial mode to 1e-7; the serial port otherwise), fft_serial.cpp:141-261
    step()
    torch.cuda.synchronize()
    parity = None
    if parity_mode != "none":
        mine = d_out[:, :, 0].contiguous() if parity_mode == "oracle" else d_out.contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            try:   # rank-0-only work: a failure here must not skip the collectives below (the other ranks wait in them)
                parity = sharded_parity(fdr, torch, torch.cat(parts)[:H].cpu().numpy(), parity_mode, cfg_idx, H, W, plen, pang, seed,
                                        dev, local_rank, sh)
            except Exception as e:
                parity = {"error": ("%s: %s" % (type(e).__name__, e))[:300]}
        dist.barrier()
    timed_out = back.sync_timed_out(sh)
    launches = back.last_launch_count()
    half = back.half_plane
    back_staged = bool(getattr(back, "staged", False))
    Rp, Cp = back.padded_rows, back.padded_cols
    peer_sync, native = bool(drv.peer_sync), bool(drv.native)
    drv.close()   # collective: every rank unmaps its peers' slabs before any rank frees its own
    peak, peak_src = measured_peak_gbs()
    planes_c = 1.5 if half else 2.0                                   # complex planes through the exchange and the column phase
    col_bytes_px = 56.0 if Rp >= 8192 else 24.0                       # K x 2048 block scheme: three sweeps
    bytes_p2 = col_bytes_px * Rp * (Cp / world) * planes_c            # per rank
    nvlink_bytes_phase = 8.0 * Rp * Cp / world * (world - 1) / world * planes_c   # per rank, per exchange
    staged = back_staged
    names = ["phase1_rows_fwd", "exchange1_push", "phase2_cols_wiener", "exchange3_push", "phase3_rows_inv", "phase4_pack"]
    t_x1 = ph[1] if staged else ph[0]   # fused form: the exchange is inside the row pass
    t_x3 = ph[3] if staged else ph[4]
    chan_px = H * W * 3
    res = {
        "workload": "rgb16384" if H == 16384 else "%dx%dx3" % (H, W), "image": [H, W, 3], "n_gpus": world,
        "ms_per_step": ms_step, "value": H * W / (ms_step * 1e-3) / 1e6, "unit": "Mpixel/s", "steps": steps, "warmup": warmup,
        "scaling": "strong", "half_plane": bool(half), "staged_exchanges": staged, "peer_sync": peer_sync,
        "driver": "fdr_shard_restore_rows (native, pipelined over the colour planes)" if (peer_sync and native) else "python unit pipeline",
        "phases_ms_serial_schedule": dict(zip(names + ["total"], ph)),
        "phase2_hbm": {"bytes_per_gpu": bytes_p2, "GBps": bytes_p2 / (ph[2] * 1e-3) / 1e9 if ph[2] > 0 else None,
                       "frac_of_peak": bytes_p2 / (ph[2] * 1e-3) / 1e9 / peak if ph[2] > 0 else None, "peak": peak, "peak_source": peak_src},
        "nvlink": {"bytes_per_gpu_per_exchange": nvlink_bytes_phase, "measured_peer_GBps": 770.0,
                   "bound_ms_both_exchanges": 2 * nvlink_bytes_phase / 770e9 * 1e3,
                   "exchange1_GBps": nvlink_bytes_phase / (t_x1 * 1e-3) / 1e9 if t_x1 > 0 else None,
                   "exchange3_GBps": nvlink_bytes_phase / (t_x3 * 1e-3) / 1e9 if t_x3 > 0 else None},
        "contract53": {"bytes_per_channel_pixel": CONTRACT_BYTES_PER_CHANNEL_PIXEL,
                       "GBps_aggregate": CONTRACT_BYTES_PER_CHANNEL_PIXEL * chan_px / (ms_step * 1e-3) / 1e9,
                       "frac_of_aggregate_peak_contract53": CONTRACT_BYTES_PER_CHANNEL_PIXEL * chan_px / (ms_step * 1e-3) / 1e9 / (peak * world),
                       "target_frac": 0.60, "target_ms": CONTRACT_BYTES_PER_CHANNEL_PIXEL * chan_px / (0.60 * peak * world * 1e9) * 1e3},
        "e2e": e2e, "gpu_launches_per_step": int(launches), "clocks": clocks, "parity": parity, "barrier_timed_out": bool(timed_out),
    }
    if timed_out:
        res["invalid"] = "a cross-rank barrier gave up after its 20 s time-out (a peer never arrived): the numbers above are not a measurement"
    return res


def run_sharded(args, fdr, torch, dist, world, rank, local_rank, dev, cfg_idx, H, W, plen, pang, seed):
    """--workload rgb16384 (or any single-image workload) at N > 1: the sharded measurement as the line itself."""
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    r = sharded_measure(args, fdr, torch, dist, world, rank, local_rank, dev, cfg_idx, H, W, plen, pang, seed, args.steps, args.warmup,
                        want_e2e=not args.no_e2e, parity_mode="none" if args.no_check else args.sharded_parity)
    if rank != 0:
        leave_group(dist, world)
        return 0
    line = {
        "metric": "Mpixel/s deblurred (FFT->Wiener->IFFT->normalise->8-bit pack)",
        "value": r["value"], "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "image": [H, W, 3], "psf": [plen, pang], "K": K_WIENER,
                   "parallelism": "row-sharded over %d GPUs, transposes fused as NVLink peer stores/loads, barriers + min/max over "
                                  "peer-memory flags" % world,
                   "l2": "working set larger than L2"},
        "e2e": r["e2e"], "gpu_launches": int(r["gpu_launches_per_step"] * args.steps), "clocks": r["clocks"],
        "roofline": {"bound": "hbm", "kernel": "phase2_cols_wiener", "achieved": r["phase2_hbm"]["GBps"], "peak": r["phase2_hbm"]["peak"],
                     "unit": "GB/s", "frac": r["phase2_hbm"]["frac_of_peak"], "traffic": None,
                     "peak_source": r["phase2_hbm"]["peak_source"] + " (of measured), per GPU",
                     "phases_ms": r["phases_ms_serial_schedule"], "nvlink": r["nvlink"], "pipeline": r["contract53"]},
        "cpu_baseline": None, "parity": r["parity"], "sharded": r,
    }
    emit(line)
    leave_group(dist, world)
    return 0


def leave_group(dist, world):
    """The line is out (or this rank has none to print): tear the process group down, but never wait for it -- after a failed
    sharded leg a peer may be gone, and an NCCL teardown that waits for it would keep the launcher alive."""
    if world > 1:
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        try:
            dist.destroy_process_group()
        except Exception:
            pass
        t.cancel()


def run_sharded_child(args, fdr, torch, dist, world, rank, local_rank, dev):
    """--sharded-child PATH: the row-sharded 16384^2 leg of the batch line in its OWN process group (one child per rank,
    spawned by sharded_leg_isolated below).  Rank 0 writes the result object to PATH; nothing goes to stdout."""
    try:   # never outlive the rank that spawned this child (PR_SET_PDEATHSIG), nor the time that rank allows it
        import ctypes
        import signal
        ctypes.CDLL(None).prctl(1, int(signal.SIGKILL))
    except Exception:
        pass
    limit = threading.Timer(args.sharded_timeout + 30.0, lambda: os._exit(3))
    limit.daemon = True
    limit.start()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    dist.barrier()   # the children's own process group works on this box: from here on a failure is the leg's, not the isolation's
    torch.cuda.synchronize()
    open("%s.started%d" % (args.sharded_child, rank), "w").close()
    c4, _, H4, W4, pl4, pa4 = WORKLOADS["rgb16384"]
    try:
        res = sharded_measure(args, fdr, torch, dist, world, rank, local_rank, dev, c4, H4, W4, pl4, pa4, 0xF17E0000 + c4,
                              args.sharded_steps, max(3, args.warmup), want_e2e=not args.no_e2e,
                              parity_mode="none" if args.no_check else args.sharded_parity)
    except Exception as e:
        res = {"unavailable": ("%s: %s" % (type(e).__name__, e))[:300]}
    if rank == 0:
        tmp = args.sharded_child + ".tmp"
        with open(tmp, "w") as f:
            json.dump(res, f)
        os.replace(tmp, args.sharded_child)
    leave_group(dist, world)
    return 0


def sharded_leg_isolated(args, dist, world, rank):
    """The row-sharded leg as child processes with their own process group (every rank of this job spawns one child on its
    GPU and waits for it): a crash, a hang or a CUDA error in that leg can then not take the batch measurement -- the line
    rank 0 is about to print -- with it.  Collective over the parent group (one broadcast).  Returns (result on rank 0 or
    None, whether this rank's child got as far as a working process group of its own)."""
    import socket
    import tempfile
    info = [None]
    if rank == 0:
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
        sk.close()
        fd, path = tempfile.mkstemp(prefix="fdr_sharded_", suffix=".json")
        os.close(fd)
        os.unlink(path)
        info = [(port, path)]
    dist.broadcast_object_list(info, src=0)
    port, path = info[0]
    # the children rendezvous among themselves: no elastic-agent store, a fresh port (child rank 0 hosts the store)
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}
    env["MASTER_ADDR"] = "127.0.0.1"
    env["MASTER_PORT"] = str(port)
    cmd = [sys.executable, CHILD_ENTRY, "--sharded-child", path, "--gpus", str(world), "--warmup", str(args.warmup),
           "--sharded-steps", str(args.sharded_steps), "--e2e-steps", str(args.e2e_steps), "--sharded-parity", args.sharded_parity,
           "--sharded-timeout", str(args.sharded_timeout)]
    cmd += (["--no-e2e"] if args.no_e2e else []) + (["--no-check"] if args.no_check else [])
    t0 = time.perf_counter()
    why = None
    try:
        p = subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL)
        try:
            rc = p.wait(timeout=args.sharded_timeout)
            if rc != 0:
                why = "child of rank %d exited with code %s" % (rank, rc)
        except subprocess.TimeoutExpired:
            p.kill()
            p.wait()
            why = "did not finish within %d s" % args.sharded_timeout
    except Exception as e:
        why = "could not start the child: %s" % e
    marker = "%s.started%d" % (path, rank)
    started = os.path.exists(marker)
    if started:
        os.unlink(marker)
    if rank != 0:
        return None, started
    res = None
    try:
        if os.path.exists(path):
            with open(path) as f:
                res = json.load(f)
            os.unlink(path)
    except Exception as e:
        why = "unreadable result: %s" % e
    if not isinstance(res, dict):
        res = {"unavailable": "row-sharded leg (child processes): %s" % (why or "no result written")}
    res["isolation"] = ("own process group, one child process per rank on the same GPUs (a failure in this leg cannot take the batch "
                        "measurement with it); %.0f s including start-up" % (time.perf_counter() - t0))
    return res, started


def single_image_lines(timeout_s=120):
    """The single-image configurations of BASELINE.json (configs[2] 4096^2, configs[4] 16384^2 on ONE GPU, configs[1] car,
    configs[0] cat geometry) measured by this same script in child processes, so that the line the driver collects carries them
    with their own clock records: `python bench.py --workload X [--flush-l2] ...` each, trimmed to the numbers.  Small images
    flush L2 between steps and time every step on its own; 16384^2 skips the oracle check here (minutes of CPU; the GPU test
    suite and the multi-GPU `sharded.parity` hold it)."""
    runs = (("rgb4096", ["--flush-l2", "--steps", "20"]), ("rgb16384", ["--steps", "10", "--no-check"]),
            ("car", ["--flush-l2", "--steps", "30"]), ("cat", ["--flush-l2", "--steps", "30"]))
    out = {}
    for wl, extra in runs:
        cmd = [sys.executable, CHILD_ENTRY, "--workload", wl, "--warmup", "3", "--no-cpu-baseline", "--no-side", "--no-e2e"] + extra
        t0 = time.perf_counter()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not last:
                out[wl] = {"unavailable": "rc=%s %s" % (r.returncode, (r.stderr or "")[-200:])}
                continue
            d = json.loads(last[-1])
            rf = d.get("roofline") or {}
            out[wl] = {"image": d["config"]["image"], "psf": d["config"]["psf"], "l2": d["config"]["l2"],
                       "ms_per_step": d["ms_per_step"], "value": d["value"], "unit": d["unit"], "steps": d["steps"],
                       "gpu_launches_per_step": d["gpu_launches"] / max(1, d["steps"]), "clocks": d["clocks"], "parity": d["parity"],
                       "kernels_in_step": (rf.get("in_step") or {}).get("kernels"), "pipeline": rf.get("pipeline"),
                       "command": "bench.py " + " ".join(cmd[2:]), "wall_s": round(time.perf_counter() - t0, 1)}
        except Exception as e:
            out[wl] = {"unavailable": ("%s: %s" % (type(e).__name__, e))[:200]}
    return out


def sample_image_comparisons():
    """BASELINE configs[1] / configs[0] as the reference states them -- `./gpu input/car_blurred.png 40 45`, `... cat_blurred.png
    50 30` -- on the repo's two sample images: this repository's CLI (the reference's own timing lines, gpu.cpp:96-113) beside
    the reference's unmodified gpu mode (fft_gpu.cu via oracle/_ref/libref_gpu.so) and its openmp / serial modes on the same
    geometry, all in this run.  Child processes; side comparisons only."""
    import re
    out = {}
    exe = os.path.join(PKG, "gpu")
    O = _load("fdr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    for name, wl in (("car_blurred.png", "car"), ("cat_blurred.png", "cat")):
        cfg_idx, _, H, W, plen, pang = WORKLOADS[wl]
        png = os.path.join(ROOT, "tests", "golden", "input", name)
        e = {"image": [H, W, 3], "psf": [plen, pang]}
        try:
            r = subprocess.run([exe, png, str(plen), str(pang)], capture_output=True, text=True, timeout=120)
            m1 = re.search(r"took\(gpu\[optimize\]\): ([0-9.eE+-]+) ms", r.stdout)
            m2 = re.search(r"took\(gpu\): ([0-9.eE+-]+) ms", r.stdout)
            if r.returncode == 0 and m1:
                ms = float(m1.group(1))
                e["cli"] = {"command": "./gpu tests/golden/input/%s %d %g" % (name, plen, pang), "gpu_optimize_ms": ms,
                            "gpu_naive_ms": float(m2.group(1)) if m2 else None, "Mpixel/s": H * W / (ms * 1e-3) / 1e6,
                            "what": "wall clock of fft_gpu::wienerDeblur_RGB_optimized as the CLI prints it (3 f32 host planes in and "
                                    "out: plan, PSF spectrum, H2D, four passes, D2H inside the timed call), second call"}
            else:
                e["cli"] = {"unavailable": "rc=%s %s" % (r.returncode, (r.stderr or r.stdout)[-200:])}
        except Exception as ex:
            e["cli"] = {"unavailable": str(ex)[:200]}
        try:
            if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")):
                r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "side_refgpu.py"), str(H), str(W), str(plen), str(pang)],
                                   capture_output=True, text=True, timeout=180)
                last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                e["reference_gpu_mode"] = json.loads(last[-1]) if (r.returncode == 0 and last) else \
                    {"unavailable": "rc=%d %s" % (r.returncode, (r.stderr or r.stdout)[-200:])}
        except Exception as ex:
            e["reference_gpu_mode"] = {"unavailable": str(ex)[:200]}
        try:
            if O.have_ref():
                psf = O.port().motion_psf(plen, pang)
                threads = os.cpu_count() or 1
                cpu_restore_images(O, "openmp", cfg_idx, 0, 1, H, W, psf, threads)   # first call starts the thread pool
                t_omp = cpu_restore_images(O, "openmp", cfg_idx, 0, 1, H, W, psf, threads)
                t_ser = cpu_restore_images(O, "serial", cfg_idx, 0, 1, H, W, psf, 1)
                e["reference_cpu"] = {"openmp_ms": t_omp * 1e3, "openmp_threads": threads, "serial_ms": t_ser * 1e3,
                                      "what": "reference fft_openmp.cpp / fft_serial.cpp compiled unmodified, 3 planes of this geometry"}
        except Exception as ex:
            e["reference_cpu"] = {"unavailable": str(ex)[:200]}
        out[name] = e
    return out


def side_comparisons(fdr, torch, plan, d_in, d_out, stream, H, W, plen, pang, our_value):
    """Same B200, same run: (1) the reference's own gpu mode (fft/fft_gpu.cu compiled unmodified for sm_100a,
    oracle/_ref/libref_gpu.so) through its 3-plane host boundary as gpu.cpp:96-105 times it, beside this library through the
    same boundary; (2) an eager torch.fft (cuFFT) pipeline with the same two-planes-per-transform packing, device resident.
    Side comparisons only: neither is on the product path."""
    out = {}
    lp = os.path.join(ROOT, "oracle", "_ref", "libref_gpu.so")
    try:
        if os.path.exists(lp):
            # own process: the reference exits on any CUDA error (CHECK_CUDA, fft_gpu.cu:59-66)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "side_refgpu.py"), str(H), str(W), str(plen), str(pang)],
                               capture_output=True, text=True, timeout=180)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and last:
                out["reference_gpu_mode"] = json.loads(last[-1])
            else:
                out["reference_gpu_mode"] = {"unavailable": "rc=%d %s" % (r.returncode, (r.stderr or r.stdout)[-300:])}
    except Exception as e:
        out["reference_gpu_mode"] = {"unavailable": str(e)[:200]}
    try:
        Bc = min(64, d_in.shape[0])
        dev = d_in.device
        wf = torch.from_numpy(plan.get_wiener()).to(dev)

        def cufft_pipeline(chunk=8):
            for b0 in range(0, Bc, chunk):
                x = d_in[b0:b0 + chunk].permute(0, 3, 1, 2).to(torch.float32) * (1.0 / 255.0)
                zs = torch.cat([torch.complex(x[:, 0:1], x[:, 1:2]), torch.complex(x[:, 2:3], torch.zeros_like(x[:, 2:3]))], 1)
                f = torch.fft.ifft2(torch.fft.fft2(zs) * wf, norm="forward")
                pl = torch.cat([f[:, :1].real, f[:, :1].imag, f[:, 1:2].real], 1)
                mn = pl.amin(dim=(2, 3), keepdim=True)
                mx = pl.amax(dim=(2, 3), keepdim=True)
                n = (pl - mn) / (mx - mn)
                d_out[b0:b0 + chunk] = torch.clamp(torch.round(n * 255.0), 0, 255).to(torch.uint8).permute(0, 2, 3, 1)

        for _ in range(2):
            cufft_pipeline()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            cufft_pipeline()
        e1.record(stream)
        torch.cuda.synchronize()
        cu = e0.elapsed_time(e1) / 3
        mpx = Bc * H * W / (cu * 1e-3) / 1e6
        out["cufft_torch_fft"] = {"Mpixel/s": mpx, "ms_per_%d_images" % Bc: cu, "this_library_Mpixel/s": our_value,
                                  "speedup": our_value / mpx,
                                  "what": "eager torch.fft.fft2 / ifft2 (cuFFT) + torch elementwise ops, device resident, same packing of two "
                                          "planes per complex transform, %d images of %dx%dx3" % (Bc, H, W)}
        del wf
    except Exception as e:
        out["cufft_torch_fft"] = {"unavailable": str(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch256x2048", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="images per rank per step (0 = the workload's own count)")
    ap.add_argument("--chunk-images", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-images", type=int, default=128, help="images per rank per end-to-end step (bounds pinned host memory)")
    ap.add_argument("--ref-images", type=int, default=2, help="--impl reference: images per step")
    ap.add_argument("--ref-mode", default="openmp", choices=["openmp", "serial", "simd", "mpi"],
                    help="--impl reference: which of the reference's CPU modes to time (default: openmp, all host threads)")
    ap.add_argument("--cpu-sample-images", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flush-l2", action="store_true", help="overwrite a 512 MB scratch between steps (small workloads)")
    ap.add_argument("--replicas", action="store_true", help="single-image workloads at N>1: independent replicas instead of row sharding")
    ap.add_argument("--sharded-parity", default="oracle", choices=["oracle", "self", "none"],
                    help="row-sharded leg: check plane 0 of the whole image against the reference CPU code (default), the whole image "
                         "against this library's single-GPU path, or nothing")
    ap.add_argument("--no-sharded", action="store_true", help="N>1, batch workload: skip the row-sharded 16384^2 leg (BASELINE configs[4])")
    ap.add_argument("--sharded-steps", type=int, default=20)
    ap.add_argument("--sharded-timeout", type=int, default=240, help="seconds after which the row-sharded leg is abandoned (the batch line is still printed)")
    ap.add_argument("--sharded-inprocess", action="store_true", help="run the row-sharded leg inside the ranks of this job instead of child processes")
    ap.add_argument("--sharded-child", default="", help=argparse.SUPPRESS)   # internal: see run_sharded_child
    ap.add_argument("--no-singles", action="store_true", help="N=1, default workload: skip the single-image configurations (child runs of this script)")
    ap.add_argument("--no-side", action="store_true", help="N=1: skip the side comparisons (reference gpu mode, cuFFT) and the extra CPU modes")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fdr = _load("fdr_b200_binding", os.path.join(PKG, "fdr.py"))

    cfg_idx, images, H, W, plen, pang = WORKLOADS[args.workload]
    B = args.images or images
    seed = 0xF17E0000 + cfg_idx
    dev = torch.device("cuda", local_rank)
    if args.sharded_child:
        return run_sharded_child(args, fdr, torch, dist, world, rank, local_rank, dev)
    if world > 1 and images == 1 and not args.replicas:
        return run_sharded(args, fdr, torch, dist, world, rank, local_rank, dev, cfg_idx, H, W, plen, pang, seed)
    stream = torch.cuda.Stream(device=dev)  # explicit non-default stream: kernels, events and copies all on it
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream

    plan = fdr.Plan(H, W, 3, max_images=B, device=local_rank)
    if args.chunk_images:
        plan.set_chunk_images(args.chunk_images)
    plan.set_psf_motion(plen, pang, K_WIENER)  # PSF built on the device
    d_in = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(d_in)
    first_image = rank * B  # each rank restores its own images (weak scaling, no collective)
    fdr.synth_images_device_u8(d_in.data_ptr(), seed, first_image, B, 3, H, W, sh)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None
    torch.cuda.synchronize()

    def step():
        if flush is not None:
            fdr.l2_flush(flush.data_ptr(), flush.numel(), sh)
        plan.restore_images_device_u8(d_in.data_ptr(), d_out.data_ptr(), B, sh)

    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi takes ~1 s to start: begin before the warm-up so the timed region is covered
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t_w0 = time.perf_counter()
    while len(sampler.lines) < 1 and time.perf_counter() - t_w0 < 4.0:   # keep the GPU under this load until the sampler reports
        step()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: device-resident throughput, CUDA events on the launching stream ----
    # Large chunks: the library's per-kernel events stay on inside the timed region (they feed roofline.in_step and cost
    # nothing measurable).  Single small images (--flush-l2 path): those events cost ~20 us of a 35-90 us step, so the timed
    # steps run without them and the per-kernel breakdown comes from a second, untimed set of steps.
    ktiming_in_region = (flush is None) and os.environ.get("FDR_BENCH_NO_KTIMING", "0") != "1"
    plan.set_kernel_timing(ktiming_in_region)
    barrier()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
    else:
        # small workloads: flush L2 between steps and time each step on its own (flush not counted)
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a0, a1 in pairs:
            fdr.l2_flush(flush.data_ptr(), flush.numel(), sh)
            a0.record(stream)
            plan.restore_images_device_u8(d_in.data_ptr(), d_out.data_ptr(), B, sh)
            a1.record(stream)
        barrier()
        ms = sum(a0.elapsed_time(a1) for a0, a1 in pairs)
        plan.set_kernel_timing(True)   # breakdown pass (not part of `value`)
        for _ in range(args.steps):
            fdr.l2_flush(flush.data_ptr(), flush.numel(), sh)
            plan.restore_images_device_u8(d_in.data_ptr(), d_out.data_ptr(), B, sh)
        barrier()
    launches_per_step = plan.last_launch_count()
    ktimes = plan.kernel_timing()
    plan.set_kernel_timing(False)
    t_b0 = time.perf_counter()
    while len(sampler.lines) < 3 and time.perf_counter() - t_b0 < 1.5:   # same load a little longer: the sampler ticks every 200 ms
        step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    # the same kernels timed ALONE (no other chunk in flight), CUDA events inside the library
    l2 = ktimes["pass2_cols_wiener"]["launches"] / max(1, args.steps)  # pass-2 launches per step = chunks per step
    iso_pairs = max(1, int(round((B * 3 / 2.0) / max(1.0, l2))))           # plane pairs per launch inside the step
    isolated = {"pairs_per_launch": iso_pairs,
                "pass1_ms": plan.time_pass(1, 0, iso_pairs, 10), "pass2_ms": plan.time_pass(2, 0, iso_pairs, 10),
                "pass3_ms": plan.time_pass(3, 0, iso_pairs, 10)}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    px_per_step = B * H * W * world
    value = px_per_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways) ----
    e2e = None
    if not args.no_e2e:
        # bounded pinned sample (<= 1.6 GB each way per rank) so that 8 ranks fit any host
        Be = min(B, max(1, args.e2e_images))
        hin = fdr.PinnedArray((Be, H, W, 3), np.uint8)
        hout = fdr.PinnedArray((Be, H, W, 3), np.uint8)
        d_in_cpu = d_in[:Be].cpu().numpy()
        np.copyto(hin.array, d_in_cpu)
        del d_in_cpu
        plan.restore_images_u8(hin.array, hout.array)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            plan.restore_images_u8(hin.array, hout.array)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": Be * H * W * world * args.e2e_steps / float(tt.item()) / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": int(hin.nbytes) * world, "d2h_bytes_per_step": int(hout.nbytes) * world,
               "images_per_gpu_per_step": Be, "steps": args.e2e_steps, "ms_per_step": float(tt.item()) / args.e2e_steps * 1e3,
               "api": "fdr_restore_images_host_u8 (pinned host in -> pinned host out, H2D/compute/D2H pipelined)"}
        e2e_first = hout.array[0].copy()
        hin.free()
        hout.free()
        pc = pcie_roofline(torch, dist, dev, world)
        bound_ms = max(e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"]) / world / (pc["GBps_each_way_per_gpu"] * 1e9) * 1e3
        e2e["roofline"] = {"bound": "pcie", "measured": pc, "bound_ms_per_step": bound_ms, "frac": bound_ms / e2e["ms_per_step"],
                           "note": "end to end is bound by the host link: pinned H2D and D2H run concurrently with the kernels; with N ranks "
                                   "on one host the copies share the host's memory and PCIe root complexes, so the per-GPU rate drops as N grows"}

    _trace('timed region + e2e done')
    padded = plan.padded
    check_idx = sorted({0, max(0, B // 2 - 1), B - 1})   # first, middle, last image of the batch
    got = {i: d_out[i].cpu().numpy() for i in check_idx} if rank == 0 else {}
    got0 = got.get(0)
    side = None
    if rank == 0 and world == 1 and not args.no_side:
        side = side_comparisons(fdr, torch, plan, d_in, d_out, stream, H, W, plen, pang, value)
    _trace('side done')
    # ---- BASELINE configs[4] inside the same line when N > 1: one 16384^2 RGB image row-sharded over all ranks ----
    run_shard_leg = world > 1 and not args.no_sharded and args.workload == "batch256x2048"
    if run_shard_leg:
        del d_in, d_out
        plan.close()
        plan = None
        torch.cuda.empty_cache()

    def sharded_leg(line):
        """The row-sharded 16384^2 measurement.  Default: child processes with their own process group (sharded_leg_isolated).
        --sharded-inprocess: in this process under a watchdog -- whatever happens in it short of a crash (a rank failing, a
        barrier that never completes), rank 0 still prints the batch line, with the reason instead of the numbers, and every
        rank exits."""
        def abort():
            try:
                if rank == 0 and line is not None:
                    line["sharded"] = {"unavailable": "row-sharded leg did not finish within %d s" % args.sharded_timeout}
                    emit(line)
            finally:
                os._exit(0)

        def watchdog(seconds):
            t = threading.Timer(seconds, abort)
            t.daemon = True
            t.start()
            return t

        fell_back = None
        if not args.sharded_inprocess:
            wd = watchdog(args.sharded_timeout + 150)   # the children are killed at sharded_timeout; the rest is slack for the decision below
            try:
                res, started = sharded_leg_isolated(args, dist, world, rank)
            except Exception as e:
                res, started = {"unavailable": ("%s: %s" % (type(e).__name__, e))[:300]}, True
            # Children that never got a working process group of their own (this box cannot run a second set of ranks on the
            # same GPUs) say nothing about the leg itself: then, and only then, run it in process.  Decided collectively.
            ok = rank == 0 and isinstance(res, dict) and "unavailable" not in res
            vote = torch.tensor([1.0 if ok else 0.0, 0.0 if started else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(vote, op=dist.ReduceOp.MAX)
            wd.cancel()
            if vote[0].item() > 0 or vote[1].item() == 0:
                return res
            fell_back = (res or {}).get("unavailable", "children did not start")
        wd = watchdog(args.sharded_timeout)
        try:
            c4, _, H4, W4, pl4, pa4 = WORKLOADS["rgb16384"]
            res = sharded_measure(args, fdr, torch, dist, world, rank, local_rank, dev, c4, H4, W4, pl4, pa4, 0xF17E0000 + c4,
                                  args.sharded_steps, max(3, args.warmup), want_e2e=not args.no_e2e,
                                  parity_mode="none" if args.no_check else args.sharded_parity)
        except Exception as e:  # keep the batch line; the other ranks leave through their own watchdogs
            res = {"unavailable": ("%s: %s" % (type(e).__name__, e))[:300]}
        wd.cancel()
        if fell_back is not None:
            res["isolation"] = "none: the child processes could not form their own process group (%s), so the leg ran inside the ranks of the job" % fell_back
        return res

    if rank != 0:
        if run_shard_leg:
            sharded_leg(None)
        leave_group(dist, world)
        return 0

    _trace('sharded leg done')
    # ---- roofline of the dominant kernel (live CUDA-event durations of this run) ----
    peak, peak_src = measured_peak_gbs()
    dom = max(ktimes, key=lambda k: ktimes[k]["ms"])
    kd = ktimes[dom]
    in_step_GBps = kd["bytes"] / (kd["ms"] * 1e-3) / 1e9 if kd["ms"] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            per_pair = json.load(open(tp)).get(args.workload, {}).get("per_plane_pair", {}).get(dom)
            traffic = per_pair * iso_pairs if per_pair else None  # per launch, like `achieved`
        except Exception:
            traffic = None
    ksum = sum(v["ms"] for v in ktimes.values())
    form_bytes = sum(v["bytes"] for v in ktimes.values())
    chan_px = B * H * W * 3 * args.steps
    Rp, Cp = padded
    iso_bytes = {"pass1_rows_fwd": (2.0 * H * W + 8.0 * H * Cp) * iso_pairs,
                 "pass2_cols_wiener": (8.0 * H * Cp + 16.0 * Rp * Cp) * iso_pairs,
                 "pass3_rows_inv_minmax": (8.0 * Rp * Cp + 8.0 * H * W) * iso_pairs}
    iso_ms = {"pass1_rows_fwd": isolated["pass1_ms"], "pass2_cols_wiener": isolated["pass2_ms"],
              "pass3_rows_inv_minmax": isolated["pass3_ms"]}
    rk = dom if dom in iso_ms else "pass2_cols_wiener"
    iso_achieved = iso_bytes[rk] / (iso_ms[rk] * 1e-3) / 1e9
    n_dom = max(1, ktimes[rk]["launches"])
    roofline = {
        "bound": "hbm", "kernel": rk, "achieved": in_step_GBps if rk == dom else iso_achieved, "peak": peak, "unit": "GB/s",
        "frac": (in_step_GBps if rk == dom else iso_achieved) / peak,
        "traffic": traffic,
        "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of an ncu --set full capture of this kernel "
                          "(profiles/traffic.json, per plane pair) scaled to the pairs of one launch; not re-measured in this run",
        "peak_source": peak_src + " (of measured, copy figure)",
        "timing": "algorithmic bytes of one launch / mean CUDA-event duration of this kernel's %d launches INSIDE the timed region "
                  "(events on the launching stream, recorded by the library around every launch)" % n_dom,
        "bytes_per_launch": ktimes[rk]["bytes"] / n_dom, "ms_per_launch": ktimes[rk]["ms"] / n_dom,
        "isolated": {"note": "the same kernels launched alone on %d plane pairs (the chunk size of the step), mean of 10 launches, right "
                             "after the timed region (secondary: the L2 is warmer and nothing else is in flight)" % iso_pairs,
                     "frac": iso_achieved / peak, "ms_per_launch": iso_ms[rk],
                     "GBps": {k: iso_bytes[k] / (iso_ms[k] * 1e-3) / 1e9 for k in iso_ms}},
        "in_step": {
            "note": "per-launch CUDA-event durations inside the timed region; chunks run on %s concurrent streams, so a launch's "
                    "duration includes time shared with other chunks' kernels" % os.environ.get("FDR_LANES", "1"),
            "kernel_share_of_step": kd["ms"] / ksum if ksum else None,
            "dominant_GBps_per_launch": in_step_GBps,
            "dominant_GBps_share_normalised": kd["bytes"] / (ms_max * (kd["ms"] / ksum) * 1e-3) / 1e9 if ksum else None,
            "kernels": {k: {"ms_per_step_summed": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                            "GBps_per_launch": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0)} for k, v in ktimes.items()},
        },
        "pipeline": {
            "formulation_bytes_per_channel_pixel": form_bytes / chan_px,
            "contract_bytes_per_channel_pixel": CONTRACT_BYTES_PER_CHANNEL_PIXEL,
            "GBps_formulation": form_bytes / (ms_max * 1e-3) / 1e9,
            "frac_of_peak_formulation": form_bytes / (ms_max * 1e-3) / 1e9 / peak,
            "GBps_contract53": CONTRACT_BYTES_PER_CHANNEL_PIXEL * chan_px / (ms_max * 1e-3) / 1e9,
            "frac_of_peak_contract53": CONTRACT_BYTES_PER_CHANNEL_PIXEL * chan_px / (ms_max * 1e-3) / 1e9 / peak,
        },
    }

    _trace('roofline done')
    # ---- parity spot check + CPU baseline on a bounded sample (rank 0 only) ----
    O = _load("fdr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    psf = O.port().motion_psf(plen, pang)
    parity = None
    cpu_baseline = None
    want_cpu = (not args.no_cpu_baseline) and world == 1
    if want_cpu or not args.no_check:
        n_s = max(1, min(B, args.cpu_sample_images)) if want_cpu else 1
        mode = "serial"
        t_cpu = 0.0
        worst = [0, 0, 0]
        for i in range(n_s):
            img = O.synth_image_u8(cfg_idx, first_image + i, H, W)
            planes = [img[c].astype(np.float32) * np.float32(1.0 / 255.0) for c in range(3)]
            t0 = time.perf_counter()
            if O.have_ref():
                outs = [O.ref().wiener(O.pad_pow2(pl), psf, K_WIENER, mode)[:H, :W] for pl in planes]
            else:
                outs = [O.port().wiener_deblur(O.pad_pow2(pl), psf, K_WIENER)["norm"][:H, :W] for pl in planes]
            t_cpu += time.perf_counter() - t0
            if i == 0 and not args.no_check:
                want = np.stack([O.port().pack_u8(o) for o in outs], -1)
                d = np.abs(got0.astype(np.int16) - want.astype(np.int16))
                worst = [int((d == 0).sum()), int((d == 1).sum()), int((d > 1).sum())]
                parity = {"image": first_image, "pixels": int(d.size), "exact": worst[0], "off_by_1": worst[1],
                          "off_by_more": worst[2], "frac_within_1": float((d <= 1).mean()),
                          "e2e_matches_device": bool(e2e is None or np.array_equal(e2e_first, got0))}
        if want_cpu:
            cpu_baseline = {"value": n_s * H * W / t_cpu / 1e6, "unit": "Mpixel/s", "cores": 1,
                            "kind": "reference" if O.have_ref() else "port",
                            "sample": "%d image(s) of %dx%dx3, reference serial mode (fft_serial.cpp compiled unmodified), %d host cores present"
                                      % (n_s, H, W, os.cpu_count() or 0)}
        if parity is not None and O.have_ref() and len(check_idx) > 1:
            # the other two images of SURVEY 8d(iv), against the reference's openmp mode (agrees with serial to ~1e-7)
            O.ref().set_threads(os.cpu_count() or 1)
            parity["more_images"] = []
            for i in check_idx[1:]:
                img = O.synth_image_u8(cfg_idx, first_image + i, H, W)
                outs = [O.ref().wiener(O.pad_pow2(img[c].astype(np.float32) * np.float32(1.0 / 255.0)), psf, K_WIENER, "openmp")[:H, :W]
                        for c in range(3)]
                want = np.stack([O.port().pack_u8(o) for o in outs], -1)
                d = np.abs(got[i].astype(np.int16) - want.astype(np.int16))
                parity["more_images"].append({"image": first_image + i, "exact": int((d == 0).sum()), "off_by_1": int((d == 1).sum()),
                                              "off_by_more": int((d > 1).sum()), "frac_within_1": float((d <= 1).mean())})
    _trace('parity + serial baseline done')
    cpu_baselines = None
    if want_cpu and O.have_ref() and not args.no_side:
        # north_star: serial, OpenMP (and the SIMD and MPI modes) timed in the same run, one image each, on this host
        cpu_baselines = {"serial": {"Mpixel/s": cpu_baseline["value"], "cores": 1}}
        threads = os.cpu_count() or 1
        for mode in ("simd", "openmp"):
            tt = cpu_restore_images(O, mode, cfg_idx, 0, 1, H, W, psf, threads)
            cpu_baselines[mode] = {"Mpixel/s": H * W / tt / 1e6, "cores": threads if mode == "openmp" else 1}
        if O.have_ref_mpi():
            ranks = max(1, min(threads, 16))
            img = O.synth_image_u8(cfg_idx, 0, H, W)
            pls = np.stack([O.pad_pow2(img[c].astype(np.float32) * np.float32(1.0 / 255.0)) for c in range(3)])
            ms_mpi = O.ref_mpi_wiener(pls, psf, K_WIENER, ranks)[1]
            cpu_baselines["mpi"] = {"Mpixel/s": H * W / (ms_mpi * 1e-3) / 1e6, "cores": ranks,
                                    "note": "reference fft_mpi.cpp over the single-node MPI stand-in (oracle/mpi_standin)"}
        cpu_baselines["sample"] = "1 image of %dx%dx3 per mode, reference sources compiled unmodified (oracle/_ref)" % (H, W)

    _trace('cpu baselines done')
    singles = samples = None
    if world == 1 and args.workload == "batch256x2048" and not args.no_side and not args.no_singles:
        del d_in, d_out
        plan.close()
        plan = None
        torch.cuda.empty_cache()
        singles = single_image_lines()
        try:
            samples = sample_image_comparisons()
        except Exception as e:
            samples = {"unavailable": ("%s: %s" % (type(e).__name__, e))[:200]}
    _trace('single images done')
    line = {
        "metric": "Mpixel/s deblurred (FFT->Wiener->IFFT->normalise->8-bit pack)",
        "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu_per_step": B, "image": [H, W, 3], "psf": [plen, pang],
                   "K": K_WIENER, "input": "u8 BGR interleaved, counter-hash (SURVEY 8d), resident in HBM",
                   "output": "u8 BGR interleaved", "l2": "flush between steps" if flush is not None else
                   "working set (%.1f GB in + out per step) larger than L2" % (2 * B * H * W * 3 / 1e9),
                   "parallelism": "images sharded across ranks, no collective" if world > 1 else "single GPU",
                   "chunk_images": args.chunk_images or "auto"},
        "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
    }
    if cpu_baselines:
        line["cpu_baselines"] = cpu_baselines
    if side:
        line["side"] = side
    if singles:
        line["single_images"] = singles
    if samples:
        line["sample_images"] = samples
    if run_shard_leg:
        line["sharded"] = sharded_leg(line)
    emit(line)
    leave_group(dist, world)
    return 0


if __name__ == "__main__":
    sys.exit(main())
