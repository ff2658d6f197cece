"""A/B check of the long-column schemes: K x 2048 blocks (col_blocks.cu) against the 128 x 128 four-step pass
(col_split.cu) on full-size planes -- restored 8-bit images must agree to within 1 LSB almost everywhere.
python profiles/check_col_blocks.py [H W ...]"""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def child(out, H, W):
    import numpy as np
    from conftest import load_fdr
    fdr = load_fdr()
    img = np.random.default_rng(H + W).integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    with fdr.Plan(H, W, 3, 1, 0) as p:
        p.set_psf_motion(50, 30.0, 0.01)
        got = p.restore_images_u8(img)
    np.save(out, got)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    import numpy as np
    sizes = [int(v) for v in sys.argv[1:]] or [16384, 2048, 8192, 8192, 16384, 16384]
    ok = True
    for H, W in zip(sizes[0::2], sizes[1::2]):
        tmp = tempfile.mkdtemp()
        outs = {}
        for mode in ("0", "1"):
            f = os.path.join(tmp, mode + ".npy")
            o = subprocess.run([sys.executable, os.path.abspath(__file__), "child", f, str(H), str(W)], env=dict(os.environ, FDR_COL_BLOCKS=mode), capture_output=True, text=True)
            if o.returncode:
                print(o.stderr[-3000:])
                sys.exit(1)
            outs[mode] = np.load(f).astype(np.int16)
        d = np.abs(outs["0"] - outs["1"])
        print("%d x %d: %d of %d pixels differ, max |delta| %d" % (H, W, int((d > 0).sum()), d.size, int(d.max())))
        ok &= int(d.max()) <= 1 and (d > 0).mean() < 1e-3
    sys.exit(0 if ok else 1)
