#!/usr/bin/env python
"""All ranks of a Z-sharded pass emulated on ONE GPU (LocalGroup: collectives become copies), so that ncu —
which must not wrap a multi-rank command — can list every kernel a rank runs and its duration.

    python profiles/shard_emulate.py --world 8 [--shape 2048,2048,512]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from skoots_b200.sharded import LocalGroup
from skoots_b200.synthetic import make_tube_volume

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--shape", default="2048,2048,512")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
shape = tuple(int(v) for v in args.shape.split(","))
n = max(8, round(16384 * shape[0] * shape[1] * shape[2] / 2**31))
tv = make_tube_volume(shape, n, seed=0, device="cuda:0", want_mask=False, want_skeleton_dict=False)
grp = LocalGroup(shape, args.world, "cuda:0")
grp.load_volume(tv.skeleton, tv.vectors)
del tv
for _ in range(args.steps):
    out = grp.step()
torch.cuda.synchronize()
print("labelled voxels", int((out > 0).sum()), "components", int(grp.ranks[0].meta[0]))
