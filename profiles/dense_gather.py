#!/usr/bin/env python
"""The fused gather on a DENSE field (half of the voxels carry a vector), for `ncu --set full -k regex:assemble_kernel`:
the density sweep in bench.py's `extras` shows the path at 0.24 of its roofline there against 0.81 on the sparse headline
volume — every voxel goes through the per-warp work queue.

    python profiles/dense_gather.py [radius]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from skoots_b200.pipeline import assemble_instances
from skoots_b200.synthetic import make_tube_volume

radius = float(sys.argv[1]) if len(sys.argv) > 1 else 12.0
shape = (1024, 1024, 256)
tv = make_tube_volume(shape, 8000, seed=1, device="cuda:0", radius=radius, want_mask=False, want_skeleton_dict=False)
out = torch.empty(shape, dtype=torch.int32, device="cuda:0")
scale = torch.tensor((60, 60, 12))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(4):
    a.record()
    assemble_instances(tv.skeleton, tv.vectors, scale, N=1, out=out, check=False)
    b.record()
    torch.cuda.synchronize()
    print("pass %d: %.3f ms" % (k, a.elapsed_time(b)))
print("non-zero vector fraction %.3f" % float((tv.vectors != 0).any(dim=0).float().mean()))
