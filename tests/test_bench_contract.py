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


def test_traffic_json_schema():
    """bench.py multiplies profiles/traffic.json[workload]['per_plane_pair'][kernel] by the plane pairs of the launch it
    times (roofline.traffic): the four pass names must be there and be below the algorithmic bytes + 10 %."""
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["batch256x2048"]
    kinds = ("pass1_rows_fwd", "pass2_cols_wiener", "pass3_rows_inv_minmax", "pass4_normalize_pack")
    for k in kinds:
        assert t["per_plane_pair"][k] > 0
        assert t["per_plane_pair"][k] <= 1.1 * t["_algorithmic_bytes_per_plane_pair"][k]
