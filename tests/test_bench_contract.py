"""CPU test: `bench.py --impl reference` prints one JSON line with the contract's keys (tiny
workload so it runs in seconds)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "car",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"] == "car"
