"""ncu target: each pass-2 variant once on `npairs` plane pairs of n x n (see fdr_plan_time_pass)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_fdr
fdr = load_fdr()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 5, 6]
with fdr.Plan(n, n, 3) as p:
    p.set_psf_motion(50, 30.0, 0.01)
    for v in variants:
        print(v, p.time_pass(2, v, npairs, 1))
