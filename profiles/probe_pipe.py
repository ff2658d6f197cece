"""ncu target: each pass-2 variant once on 6 plane pairs of 2048 x 2048 (see fdr_plan_time_pass)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_fdr
fdr = load_fdr()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
with fdr.Plan(n, n, 3) as p:
    p.set_psf_motion(50, 30.0, 0.01)
    for v in (4, 5):
        print(v, p.time_pass(2, v, npairs, 1))
    print(1, p.time_pass(1, 0, npairs, 1))
    print(3, p.time_pass(3, 0, npairs, 1))
