"""CPU test of bench.py's multi-rank line: `python -m torch.distributed.run --nproc-per-node 2 tests/fake_gpu_bench.py ...`
launches bench.main() exactly as the driver launches bench.py, with the device layer replaced by CPU stand-ins
(tests/fake_gpu_bench.py) and gloo instead of NCCL.  What is checked is the host logic a GPU box cannot be spared for:
every rank issues the same number of restores (loop counts come from all-reduced durations), the `sharded` object and the
contract keys are in the ONE line rank 0 prints, and the watchdog keeps that line when the row-sharded leg hangs or fails."""
import json
import os
import stat
import subprocess
import sys

import pytest

from conftest import ROOT

pytest.importorskip("torch")

FAKE_SMI = """#!/bin/bash
while true; do echo "1965, 1965, 512.3, Not Active, Not Active, Not Active, Active"; sleep 0.2; done
"""


def run_fake_bench(tmp_path, port, extra_args=(), extra_env=None, nproc=2):
    bindir = tmp_path / "bin"
    bindir.mkdir(exist_ok=True)
    smi = bindir / "nvidia-smi"
    smi.write_text(FAKE_SMI)
    smi.chmod(smi.stat().st_mode | stat.S_IEXEC)
    shared = tmp_path / "shared"
    shared.mkdir(exist_ok=True)
    env = dict(os.environ)
    env["PATH"] = str(bindir) + os.pathsep + env.get("PATH", "")
    env["FAKE_TMP"] = str(shared)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "fake_gpu_bench.py"), "--gpus", str(nproc), "--steps", "2",
           "--warmup", "3"] + list(extra_args)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    return r, lines


def base_port():
    return 29800 + os.getpid() % 1500


@pytest.mark.parametrize("inprocess", [False, True])
def test_two_rank_line_carries_the_sharded_leg(tmp_path, inprocess):
    """Default: the row-sharded leg runs as child processes with their own process group; --sharded-inprocess: inside the
    ranks of the job, under the watchdog."""
    r, lines = run_fake_bench(tmp_path, base_port() + (10 if inprocess else 0), ["--sharded-inprocess"] if inprocess else [])
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1, "stdout must hold exactly ONE line: %r" % lines   # the contract
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity", "sharded"):
        assert k in d, k
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["workload"] == "batch256x2048"
    assert d["clocks"]["samples"] > 0 and d["clocks"]["reasons"] == ["sw_power_cap"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["roofline"]["bound"] == "pcie"
    assert d["parity"]["off_by_more"] == 0 and len(d["parity"]["more_images"]) >= 1
    s = d["sharded"]
    assert "unavailable" not in s, s
    assert s["n_gpus"] == 2 and s["scaling"] == "strong" and s["steps"] == 20
    assert s["clocks"]["samples"] > 0
    assert s["barrier_timed_out"] is False and "invalid" not in s
    assert s["parity"]["frac_within_1"] >= 0.999 and s["parity"]["off_by_more"] == 0   # against the reference CPU code
    assert s["e2e"]["matches_device"] is True and s["e2e"]["h2d_bytes_per_step"] > 0
    assert set(s["phases_ms_serial_schedule"]) == {"phase1_rows_fwd", "exchange1_push", "phase2_cols_wiener", "exchange3_push",
                                                   "phase3_rows_inv", "phase4_pack", "total"}
    assert s["contract53"]["target_frac"] == 0.60
    assert ("isolation" in s) == (not inprocess)


@pytest.mark.parametrize("inprocess", [False])
def test_batch_line_survives_a_hung_rank(tmp_path, inprocess):
    r, lines = run_fake_bench(tmp_path, base_port() + (11 if inprocess else 1),
                              ["--sharded-timeout", "12" if not inprocess else "6"] + (["--sharded-inprocess"] if inprocess else []),
                              {"FAKE_HANG_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["n_gpus"] == 2
    # rank 0's own watchdog, or -- the hung rank's watchdog having fired first -- the collective that lost its peer
    assert "unavailable" in d["sharded"]


@pytest.mark.parametrize("bad_rank,inprocess", [(0, False), (1, True)])
def test_failure_inside_the_sharded_leg_keeps_the_batch_line(tmp_path, bad_rank, inprocess):
    r, lines = run_fake_bench(tmp_path, base_port() + 2 + bad_rank + (10 if inprocess else 0),
                              ["--sharded-timeout", "20" if not inprocess else "6"] + (["--sharded-inprocess"] if inprocess else []),
                              {"FAKE_RAISE_RANK": str(bad_rank)})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and "unavailable" in d["sharded"]


def test_batch_line_survives_a_crashing_child(tmp_path):
    """A child that dies outright (SIGKILL itself in phase 1) -- the case the in-process form cannot survive, because the
    launcher tears every rank down when one dies."""
    # generous limit: it only runs out if the children are very slow to start (then the in-process fallback would meet the crash)
    r, lines = run_fake_bench(tmp_path, base_port() + 5, ["--sharded-timeout", "90"], {"FAKE_CRASH_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and "unavailable" in d["sharded"]


def test_single_rank_line(tmp_path):
    env = {"FDR_BENCH_TRACE": "0"}
    bindir = tmp_path / "bin"
    bindir.mkdir()
    smi = bindir / "nvidia-smi"
    smi.write_text(FAKE_SMI)
    smi.chmod(smi.stat().st_mode | stat.S_IEXEC)
    (tmp_path / "shared").mkdir()
    e = dict(os.environ, PATH=str(bindir) + os.pathsep + os.environ.get("PATH", ""), FAKE_TMP=str(tmp_path / "shared"), **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fake_gpu_bench.py"), "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 1 and "sharded" not in d
    singles = d["single_images"]   # BASELINE configs[2], [4] on one GPU, [1], [0]: child runs of the same script
    assert set(singles) == {"rgb4096", "rgb16384", "car", "cat"}
    samples = d["sample_images"]   # the CLI on the repo's two sample images beside the reference's gpu / openmp / serial modes
    assert set(samples) == {"car_blurred.png", "cat_blurred.png"}
    for e in samples.values():     # no device here: the GPU legs must say so instead of inventing numbers
        assert "unavailable" in e["cli"] and e["reference_cpu"]["serial_ms"] > 0
    for wl, r in singles.items():
        assert "unavailable" not in r, (wl, r)
        assert r["ms_per_step"] > 0 and r["value"] > 0 and r["clocks"]["samples"] > 0 and r["gpu_launches_per_step"] > 0
        assert (r["parity"] is None) == (wl == "rgb16384")
        assert r["parity"] is None or r["parity"]["off_by_more"] == 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"]
    assert d["roofline"]["kernel"] == "pass2_cols_wiener" and d["roofline"]["isolated"]["frac"] > 0
    assert {"serial", "simd", "openmp"} <= set(d.get("cpu_baselines", {"serial": 0, "simd": 0, "openmp": 0}))


def test_children_that_cannot_start_fall_back_to_the_in_process_leg(tmp_path):
    """Only when the children never got a process group of their own (an environment limit, not a failure of the leg)."""
    r, lines = run_fake_bench(tmp_path, base_port() + 6, ["--sharded-timeout", "30"], {"FAKE_CHILD_NO_START": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    s = json.loads(lines[0])["sharded"]
    assert "unavailable" not in s, s
    assert s["isolation"].startswith("none:") and s["parity"]["off_by_more"] == 0 and s["n_gpus"] == 2


def test_explicit_sharded_workload_is_the_line_itself(tmp_path):
    """`--workload rgb16384 --gpus N`: the row-sharded measurement as the contract line (strong scaling)."""
    r, lines = run_fake_bench(tmp_path, base_port() + 7, ["--workload", "rgb16384"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["scaling"] == "strong" and d["n_gpus"] == 2 and d["config"]["workload"] == "rgb16384"
    assert d["parity"]["off_by_more"] == 0 and d["clocks"]["samples"] > 0 and d["roofline"]["kernel"] == "phase2_cols_wiener"
    assert d["sharded"]["e2e"]["matches_device"] is True
