#!/usr/bin/env python
"""Runs every training-side / stencil kernel a few times on its SURVEY §8d configuration so that
`ncu --metrics gpu__time_duration.sum` can list their device durations (these kernels run for 5-100 us: timing a
single Python call with CUDA events mostly measures the host's launch path, see profiles/README.md).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small.csv python profiles/small_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from skoots_b200.lib import embedding_to_prob as e2p
from skoots_b200.lib import morphology as morph
from skoots_b200.lib import skeleton as skel
from skoots_b200.lib import vector_to_embedding as v2e
from skoots_b200.pipeline import tile_epilogue
from skoots_b200.synthetic import make_tube_volume

DEV = torch.device("cuda:0")
B = 8
vols = [make_tube_volume((300, 300, 20), 20, seed=s) for s in range(B)]
present = [{int(k): t.skeletons[int(k)].to(DEV) for k in torch.unique(t.mask).tolist() if k != 0} for t in vols]
masks = torch.stack([t.mask for t in vols]).to(DEV)
vec = torch.stack([t.vectors.float() for t in vols]).to(torch.bfloat16).to(DEV)
scale, sigma, an = torch.tensor((60.0, 60.0, 12.0)), torch.tensor((20.0, 20.0, 20.0)), (1.0, 1.0, 3.0)
g = torch.Generator().manual_seed(1)
unet = torch.rand((1, 5, 300, 300, 20), generator=g).to(DEV)
gv = torch.zeros((3, 300, 300, 20), dtype=torch.float16, device=DEV)
gs = torch.zeros((1, 300, 300, 20), dtype=torch.uint8, device=DEV)
img = unet[:, 3:4].contiguous()
for _ in range(3):
    baked = skel.bake_skeletons_batch(masks, present, an, average=True, check=False)
    raw = skel.bake_skeletons_batch(masks, present, an, average=False, check=False)
    skel.average_baked_skeletons(raw)
    bk = baked.to(torch.bfloat16)
    p = e2p.vector_to_prob(scale, vec, bk, sigma)
    emb = v2e.vector_to_embedding(scale, vec)
    e2p.baked_embed_to_prob(emb, bk, sigma)
    vg = vec.clone().requires_grad_(True)
    e2p.vector_to_prob(scale, vg, bk, sigma).sum().backward()
    tile_epilogue(unet, gv, gs, (0, 0, 0), (50, 50, 5))
    morph.binary_dilation(img)
    morph.binary_dilation_2d(img)
    morph.binary_erosion(img)
    skel.skeleton_to_mask(present[0], (300, 300, 20), radius=9, flank_radius=3)
torch.cuda.synchronize()
print("done")
