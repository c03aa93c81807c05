#!/usr/bin/env python
"""Lists, per kernel of skoots_b200/libskoots_b200.so, how many SASS instructions carry the mnemonics that show which
hardware paths the kernel uses (cuobjdump -sass on the sm_100a cubin; runs on the CPU build box):

    python profiles/sass_markers.py > profiles/r02_sass_markers.txt

UBLKCP = cp.async.bulk (the TMA engine's bulk copy), SYNCS = mbarrier operations, LDGSTS = cp.async, LDG.E.128 / STG.E.128
= 16-byte global loads / stores (.NA = L1 no-allocate, the streaming form), ATOMS / ATOMG = shared / global atomics, VOTE / SHFL / REDUX = warp votes, shuffles and
reductions, POPC / FLO = popcount / find-leading-one (run extraction on the bit mask), MEMBAR + *.STRONG.SYS = the
system-scope release / acquire of the peer-mailbox flags.  No HMMA / UTCHMMA appears anywhere: the path is integer and
byte work with no contraction, so there is nothing for the tensor cores to do.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "skoots_b200", "libskoots_b200.so")
MARKS = ["UBLKCP", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.NA.128", "STG.E.128", "STG.E.NA.128", "LDS.128", "ATOMS", "ATOMG", "VOTE", "SHFL", "REDUX", "POPC", "FLO",
         "BAR.SYNC", "MEMBAR", "STRONG.SYS", "HMMA", "UTCHMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, fn = collections.defaultdict(collections.Counter), None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    if fn and "/*" in ln:
        for k in MARKS:
            if k in ln:
                counts[fn][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(__doc__.split("\n\n")[1].strip() if False else "# SASS markers per kernel (see profiles/sass_markers.py for the legend)")
rows = []
for mangled, name in zip(counts, names):
    short = name.split("(")[0].replace("void ", "").strip() or name[:60]
    rows.append((short, counts[mangled]))
for short, c in sorted(rows):
    print(f"{short[:84]:84s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
total = collections.Counter()
for _, c in rows:
    total.update(c)
print("# totals: " + " ".join(f"{k}={v}" for k, v in sorted(total.items())) + f"  HMMA={total.get('HMMA', 0)} UTCHMMA={total.get('UTCHMMA', 0)}")
