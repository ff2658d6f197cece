"""A/B check of the pass-2 kernel variants (col_tma.cu per-tile / pipelined, col_wide.cu): restored 8-bit
images of each variant against the per-tile TMA kernel (count of differing pixels, max |delta|) and the
pass-2 time per plane pair.   python profiles/check_col_variants.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

CASES = [(2048, 8), (2048, 3), (1024, 16), (4096, 2)]
MODES = {"per-tile": {"FDR_COL_PIPE": "0", "FDR_COL_WIDE": "0"}, "pipelined": {"FDR_COL_PIPE": "1", "FDR_COL_WIDE": "0"},
         "wide": {"FDR_COL_PIPE": "0", "FDR_COL_WIDE": "1"}}


def child(outdir):
    import numpy as np
    from conftest import load_fdr
    fdr = load_fdr()
    res = {}
    for n, nimg in CASES:
        imgs = np.random.default_rng(n + nimg).integers(0, 256, (nimg, n, n, 3), dtype=np.uint8)
        with fdr.Plan(n, n, 3, nimg, 0) as p:
            p.set_psf_motion(50, 30.0, 0.01)
            got = p.restore_images_u8(imgs)
            npairs = 12 if n <= 2048 else 3
            t = {v: p.time_pass(2, v, npairs) / npairs * 1e3 for v in (4, 5, 6) if not (v == 6 and n not in (1024, 2048, 4096))}
        np.save(os.path.join(outdir, "%dx%d.npy" % (n, nimg)), got)
        res["%dx%d" % (n, nimg)] = t
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    import tempfile
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    import numpy as np
    tmp = tempfile.mkdtemp()
    times = None
    for mode, env in MODES.items():
        d = os.path.join(tmp, mode)
        os.makedirs(d)
        o = subprocess.run([sys.executable, os.path.abspath(__file__), "child", d], env=dict(os.environ, **env), capture_output=True, text=True)
        line = [l for l in o.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(o.stdout[-2000:], o.stderr[-4000:])
            sys.exit(1)
        times = json.loads(line[0][7:])
    ok = True
    for n, nimg in CASES:
        k = "%dx%d" % (n, nimg)
        base = np.load(os.path.join(tmp, "per-tile", k + ".npy")).astype(np.int16)
        msg = []
        for mode in ("pipelined", "wide"):
            d = np.abs(np.load(os.path.join(tmp, mode, k + ".npy")).astype(np.int16) - base)
            msg.append("%s: %d of %d pixels differ, max |delta| %d" % (mode, int((d > 0).sum()), d.size, int(d.max())))
            ok &= int(d.max()) <= 1 and (d > 0).mean() < 1e-3
        print("%-8s %s | us per pair: %s" % (k, "; ".join(msg), {{"4": "per-tile", "5": "pipelined", "6": "wide"}[a]: round(b, 2) for a, b in times[k].items()}))
    sys.exit(0 if ok else 1)
