#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that show what the build is made of (TMA tensor loads/stores, bulk copies, mbarrier
waits, packed fp32x2 arithmetic, byte loads, peer-sync loads/stores).  No GPU needed:
    python profiles/sass_opcodes.py > profiles/r2/sass_opcodes.txt
Reads lib/libfdr_b200.so (cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "parallel-implementation-of-frequency-domain-image-restoration-using-fft_b200", "lib", "libfdr_b200.so")
OPS = ["UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FADD2", "FFMA2", "FMUL2", "FFMA", "FADD", "FMUL", "LDG.E.U8", "LDG", "STG", "LDS", "STS",
       "BAR.SYNC", "ATOMG", "RED", "LD.E.STRONG.SYS", "ST.E.STRONG.SYS", "MEMBAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["total"] += 1
        for k in OPS:
            if op == k or op.startswith(k + ".") or (k == "LDG.E.U8" and op.startswith("LDG") and ".U8" in op):
                counts[cur][k] += 1
                break
    names = demangle(list(counts))
    tot = collections.Counter()
    print("# %d kernels in %s (sm_100a SASS); columns: instruction counts (static, per kernel)" % (len(counts), os.path.relpath(LIB, ROOT)))
    print("# UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA tiles), UBLKCP = cp.async.bulk (1-D bulk copies), UBLKPF = bulk L2 prefetch,")
    print("# SYNCS = mbarrier ops, FADD2/FFMA2/FMUL2 = packed fp32x2, LDG.E.U8 = single-byte global loads, *.STRONG.SYS = peer-sync flags")
    cols = ["total"] + OPS
    print("%-110s %s" % ("kernel", " ".join("%9s" % c[:9] for c in cols)))
    for k, c in counts.items():
        nm = re.sub(r"\s+", " ", names.get(k, k))
        nm = re.sub(r"\(.*$", "", nm.replace("(anonymous namespace)::", ""))[:108]
        print("%-110s %s" % (nm, " ".join("%9d" % c[x] for x in cols)))
        tot.update(c)
    print("%-110s %s" % ("ALL KERNELS", " ".join("%9d" % tot[x] for x in cols)))


if __name__ == "__main__":
    sys.exit(main())
