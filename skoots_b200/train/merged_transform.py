"""`skoots.train.merged_transform.elastic_deform` on B200 (reference: skoots/train/merged_transform.py:75-188, 43-72).

SURVEY.md §8 row f4, the part that can be pinned to the reference: the elastic deformation of image, mask and skeleton
points.  (The skeletonisation half of f4 — `skimage.morphology.skeletonize(method="lee")` / kimimaro,
train/generate_skeletons.py:65-157 — depends on libraries absent from this image and from the reference tree: it stays
unpinned and unbuilt.)

Same signature and return convention as the reference: `elastic_deform(*args, skeleton=..., displacement_shape=...,
displacement_magnitude=...) -> (*deformed_args, skeleton)`.  The coarse random field is drawn with the same call the
reference makes (`torch.rand(displacement_shape, device=device)`), so a seeded run consumes the generator identically;
`noise=` injects a field instead (tests).  The two dense (X,Y,Z,3) grids of the reference are never built.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L


def elastic_deform(*args: Tensor, skeleton: Dict[int, Tensor], displacement_shape: Tuple[int, int, int] = (6, 6, 2),
                   displacement_magnitude: Tuple[float, float, float] = (0.05, 0.05, 0.01), noise: Optional[Tensor] = None):
    assert len(args) > 0, "must pass at least one positional argument"
    shape = args[0].shape
    for i, a in enumerate(args):
        assert isinstance(a, Tensor), f"positional argument {i} must be of type torch.Tensor not {type(a)}"
        assert a.shape == shape, f"positional argument {i} must be of {shape=} not {a.shape}"
    assert args[0].ndim == 5, f"image must be in shape: [B, C, X, Y, Z], not {args[0].shape}"
    assert len(displacement_shape) == 3, "displacement_shape must be a tuple of integers with len == 3"
    assert len(displacement_magnitude) == 3, "displacement_magnitude must be a tuple of integers with len == 3"
    assert max(displacement_magnitude) < 1.0, "max displacement must not exceed 1.0"
    dev = L.require_cuda(*args)
    b, c, x, y, z = shape
    dshape = (1, 3, displacement_shape[2], displacement_shape[1], displacement_shape[0])  # merged_transform.py:129-135
    mag = tuple(reversed(displacement_magnitude))                                          # :136
    if noise is None:
        noise = torch.rand(dshape, device=dev)                                             # :140
    noise = noise.to(device=dev, dtype=torch.float32).reshape(dshape).contiguous()
    lib = L.load()
    out = []
    with torch.cuda.device(dev):
        for a in args:
            src = a.float().contiguous()
            dst = torch.empty_like(src)
            L.check(lib.skb_elastic_resample(noise.data_ptr(), dshape[2], dshape[3], dshape[4], L.f3(mag), src.data_ptr(),
                                             dst.data_ptr(), b * c, x, y, z, L.stream_ptr(dev)))
            out.append(dst)
        # all skeletons in one launch: concatenate, transform, split back (the reference loops over the ids, :47-72)
        keys = list(skeleton.keys())
        if keys:
            vals = [skeleton[k] for k in keys]
            is_int = not vals[0].is_floating_point()
            cat = torch.cat([v.to(dev).reshape(-1, 3) for v in vals], 0)
            cat = (cat.to(torch.int64) if is_int else cat.float()).contiguous()
            new = torch.empty_like(cat)
            L.check(lib.skb_elastic_points(noise.data_ptr(), dshape[2], dshape[3], dshape[4], L.f3(mag), cat.data_ptr(), int(is_int),
                                           cat.shape[0], x, y, z, new.data_ptr(), L.stream_ptr(dev)))
            at = 0
            skeleton = dict(skeleton)
            for k, v in zip(keys, vals):
                n = v.reshape(-1, 3).shape[0]
                skeleton[k] = new[at:at + n].to(v.dtype).reshape(v.shape)
                at += n
    return (*out, skeleton)
