#!/usr/bin/env python
"""Timeline of one split pass (pipeline.assemble_split) from CUDA events on its two streams, next to the
parts run alone: where the labelling chain sits relative to the gather's stream phase.

    python profiles/split_timeline.py [--shape 2048,2048,512] [--steps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import skoots_b200._lib as L
from skoots_b200.lib.flood_fill import launch_label, new_sparse
from skoots_b200.pipeline import assemble_split
from skoots_b200.synthetic import make_tube_volume

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="2048,2048,512")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
shape = tuple(int(v) for v in args.shape.split(","))
X, Y, Z = shape
dev = torch.device("cuda:0")
n = max(8, round(16384 * X * Y * Z / 2**31))
tv = make_tube_volume(shape, n, seed=0, device=dev, want_mask=False, want_skeleton_dict=False)
mask, vec = tv.skeleton, tv.vectors
sparse = new_sparse(shape, dev)
out = torch.empty(shape, dtype=torch.int32, device=dev)
flags = torch.empty(X * Y * Z // 256, dtype=torch.int32, device=dev)
scale = (60, 60, 12)
lib = L.load()
from skoots_b200.pipeline import stream_ctas_default
CTAS = stream_ctas_default()
print("stream phase CTAs per SM:", CTAS or "non-persistent")


def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for k in range(reps + 1):
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b) if k else 0.0
    return tot / reps


for _ in range(3):
    assemble_split(mask, vec, scale, sparse, out, group_flags=flags)
torch.cuda.synchronize()
acc = {}
for _ in range(args.steps):
    tr = {}
    assemble_split(mask, vec, scale, sparse, out, group_flags=flags, trace=tr)
    torch.cuda.synchronize()
    for k, e in tr.items():
        acc[k] = acc.get(k, 0.0) + tr["begin"].elapsed_time(e) / args.steps
print("split pass, ms from its start:", {k: round(v, 3) for k, v in acc.items()})
s = L.stream_ptr(dev)
print("alone: pack         %.3f ms" % timed(lambda: launch_label(mask, sparse, False, 2, L.CCL_PHASE_PACK), args.steps))
print("alone: pack + label chain  %.3f ms" % timed(lambda: launch_label(mask, sparse, False, 2, 0), args.steps))
print("alone: stream phase %.3f ms" % timed(lambda: L.check(lib.skb_assemble_stream(
    vec.data_ptr(), L.dtype_code(vec), X, Y, Z, 0, Z, sparse.workspace.data_ptr(), flags.data_ptr(), out.data_ptr(),
    L.dtype_code(out), CTAS, s)), args.steps))
print("alone: resolve      %.3f ms" % timed(lambda: L.check(lib.skb_assemble_resolve(
    vec.data_ptr(), L.dtype_code(vec), X, Y, Z, 0, Z, L.f3(scale), sparse.workspace.data_ptr(), 0, 0, flags.data_ptr(),
    out.data_ptr(), L.dtype_code(out), s)), args.steps))
print("flagged groups: %.2f %% of %d" % (100.0 * sum(bin(v & 0xffffffff).count("1") for v in flags[:200000].tolist()) / (32 * 200000), X * Y * Z // 8))
