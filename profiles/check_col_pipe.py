"""A/B check of the persistent pipelined column kernel (col_tma.cu) against the one-tile-per-CTA
TMA kernel: bit-identical restored images (sha256 of the u8 output) and per-pair pass-2 time.
python profiles/check_col_pipe.py            -- runs itself twice (FDR_COL_PIPE=0 / 1) and compares."""
import hashlib, json, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

CASES = [(2048, 8), (4096, 2), (1024, 16), (2048, 3)]


def child():
    import numpy as np
    from conftest import load_fdr
    fdr = load_fdr()
    out = {}
    for n, nimg in CASES:
        imgs = np.random.default_rng(n + nimg).integers(0, 256, (nimg, n, n, 3), dtype=np.uint8)
        with fdr.Plan(n, n, 3, nimg, 0) as p:
            p.set_psf_motion(50, 30.0, 0.01)
            got = p.restore_images_u8(imgs)
            t = {v: p.time_pass(2, v, 6 if n <= 2048 else 2) for v in (4, 5)}
        out["%dx%d" % (n, nimg)] = {"sha": hashlib.sha256(got.tobytes()).hexdigest(), "ms": t}
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for mode in ("0", "1"):
        env = dict(os.environ, FDR_COL_PIPE=mode)
        o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        line = [l for l in o.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(o.stdout[-2000:], o.stderr[-4000:]); sys.exit(1)
        res[mode] = json.loads(line[0][7:])
    ok = True
    for k in res["0"]:
        same = res["0"][k]["sha"] == res["1"][k]["sha"]
        ok &= same
        ms = res["1"][k]["ms"]
        print("%-10s identical=%s  pass 2 per launch: per-tile TMA %.1f us, pipelined %.1f us" % (k, same, ms["4"] * 1e3, ms["5"] * 1e3))
    sys.exit(0 if ok else 1)
