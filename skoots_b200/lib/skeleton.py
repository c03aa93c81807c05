"""`skoots.lib.skeleton` on B200 (reference: skoots/lib/skeleton.py)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _lib as L


def index_skeleton_by_embed(skeleton: Tensor, embed: Tensor) -> Tensor:
    """skeleton (1,1,Xs,Ys,Zs) labels, embed (1,3,x,y,z) fp32 -> (1,1,x,y,z) int32
    (skeleton.py:656-695: rint, clamp to the label volume, gather)."""
    assert embed.device == skeleton.device, "embed and skeleton must be on same device"
    assert (
        embed.ndim == 5 and skeleton.ndim == 5
    ), "Embed and skeleton must be a 5D tensor"
    dev = L.require_cuda(skeleton, embed)
    b, c, x, y, z = embed.shape
    if b != 1 or c != 3:
        raise RuntimeError(f"embed must have shape (1,3,x,y,z), got {tuple(embed.shape)}")
    if skeleton.dtype not in (torch.int16, torch.int32, torch.uint8):
        skeleton = skeleton.to(torch.int32)
    skeleton = skeleton.contiguous()
    embed = embed.float().contiguous()
    out = torch.empty((1, 1, x, y, z), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().skb_index_by_embed(skeleton.data_ptr(), L.dtype_code(skeleton), skeleton.shape[2],
                                            skeleton.shape[3], skeleton.shape[4], embed.data_ptr(), x * y * z,
                                            out.data_ptr(), L.stream_ptr(dev)))
    return out
