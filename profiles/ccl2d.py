#!/usr/bin/env python
"""Per-slice (2-D mode) CCL of a 4096x4096 stack, BASELINE.json configs[4]: label + dense write, for ncu.

    python profiles/ccl2d.py --slices 64
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from skoots_b200.lib import flood_fill as ff
from skoots_b200.synthetic import make_tube_volume

ap = argparse.ArgumentParser()
ap.add_argument("--slices", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
S = args.slices
dev = torch.device("cuda:0")
tv = make_tube_volume((min(S, 8), 4096, 4096), 4096 * min(S, 8), seed=0, device=dev, want_mask=False, want_skeleton_dict=False)
stack = tv.skeleton if S <= 8 else tv.skeleton.repeat(S // 8, 1, 1).contiguous()
del tv
sp = ff.label_components(stack, planar=True, label_base=0)
out = torch.empty(stack.shape, dtype=torch.int32, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(args.steps + 1):
    a.record()
    s2 = ff.label_components(stack, planar=True, label_base=0, workspace=sp.workspace, check=False)
    ff.write_dense(s2, out)
    b.record()
    torch.cuda.synchronize()
    if k:
        print("pass %d: %.3f ms" % (k, a.elapsed_time(b)))
print("foreground %.2f %%, components %d" % (100.0 * float((stack > 0).float().mean()), s2.num_components))
