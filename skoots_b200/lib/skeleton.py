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
    dev, staged = L.compute_device(skeleton, embed)
    if staged:  # eval() hands over the same whole-image label volume for every crop (eval.py:277-279): uploaded once
        return index_skeleton_by_embed(L.stage_in(skeleton, dev, reuse=True), L.stage_in(embed, dev)).cpu()
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


# --------------------------------------------------------------------------------------------
# training-side ops
# --------------------------------------------------------------------------------------------
from typing import Dict, Tuple  # noqa: E402

import numpy as np  # noqa: E402

_DISK_CACHE: dict = {}


def get_cached_disk_coords(device, radius: int = 7, flank_radius: int = 3) -> Tensor:
    """(3,S) stamp offsets of skoots/lib/utils.py:421-438: disk(radius) at dz=0, disk(flank) at
    dz=+-1, x/y shifted by -radius//2 (sic), in torch.nonzero order."""
    key = (str(device), int(radius), int(flank_radius))
    hit = _DISK_CACHE.get(key)
    if hit is None:
        def disk(r):
            span = np.arange(-r, r + 1)
            xx, yy = np.meshgrid(span, span)
            return (xx * xx + yy * yy) <= r * r
        centre, flank = disk(radius), disk(flank_radius)
        flank = np.pad(flank, (centre.shape[0] - flank.shape[0]) // 2)
        off = np.argwhere(np.stack((flank, centre, flank), axis=-1)).astype(np.int64)
        off[:, 2] -= 1
        off[:, :2] -= radius // 2
        hit = torch.from_numpy(off).to(device).T.contiguous()
        _DISK_CACHE[key] = hit
    return hit


def average_baked_skeletons(baked_skeleton: Tensor, kernel_size: int = 3) -> Tensor:
    """(B,3,X,Y,Z) -> per channel sum(3x3x3 window)/max(1,count(window>0)) (skeleton.py:18-48)."""
    if kernel_size != 3:
        raise NotImplementedError("the reference only ever uses kernel_size=3")
    dev, staged = L.compute_device(baked_skeleton)
    if staged:
        return average_baked_skeletons(L.stage_in(baked_skeleton, dev), kernel_size).cpu()
    src = baked_skeleton.float().contiguous()
    b, c, X, Y, Z = src.shape
    out = torch.empty_like(src)
    with torch.cuda.device(dev):
        L.check(L.load().skb_masked_mean27(src.data_ptr(), out.data_ptr(), b * c, X, Y, Z, L.stream_ptr(dev)))
    return out


def _pack_skeletons(skeletons: Dict[int, Tensor], device):
    ids = sorted(int(k) for k in skeletons.keys())
    lens = [int(skeletons[k].shape[0]) for k in ids]
    offsets = np.zeros(len(ids) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    n = int(offsets[-1])
    pts = torch.zeros((max(n, 1), 4), dtype=torch.float32, device=device)
    if n:
        pts[:n, :3] = torch.cat([skeletons[k].to(device=device, dtype=torch.float32).reshape(-1, 3) for k in ids], 0)
    return (torch.tensor(ids, dtype=torch.int32, device=device), torch.from_numpy(offsets).to(device), pts, n)


def bake_skeleton(masks: Tensor, skeletons: Dict[int, Tensor], anisotropy: Tuple[float, float, float] = (1.0, 1.0, 1.0),
                  average: bool = True, device: str = "cpu", return_distance: bool = False):
    """Drop-in for skoots.lib.skeleton.bake_skeleton (:448-528) with the CPU/torch semantics
    (anisotropy scales coordinates, first minimum wins, fp32 out — SURVEY A.5).  `device` is
    accepted and ignored like the reference's positional mix-up (:507); the work runs on
    masks.device, which must be CUDA.  return_distance=True also returns the (1,X,Y,Z) distance."""
    dev, staged = L.compute_device(masks)
    if staged:
        return L.stage_out(bake_skeleton(L.stage_in(masks, dev), skeletons, anisotropy, average, device, return_distance), True)
    if -1 in skeletons:
        x, y, z = masks.shape[-3:]
        return torch.zeros((3, x, y, z), device=dev, dtype=torch.float16)
    if masks.ndim == 4 and masks.shape[0] == 1:
        masks = masks.squeeze(0)
    assert masks.ndim == 3, f"masks must be 3d with no batch. not {masks.shape=}"
    m = masks if masks.dtype in (torch.int32, torch.int16, torch.uint8) else masks.to(torch.int32)
    m = m.contiguous()
    X, Y, Z = m.shape
    ids, offsets, pts, n_pts = _pack_skeletons(skeletons, dev)
    baked = torch.empty((3, X, Y, Z), dtype=torch.float32, device=dev)
    dist = torch.empty((1, X, Y, Z), dtype=torch.float32, device=dev) if return_distance else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().skb_bake_skeleton(m.data_ptr(), L.dtype_code(m), X, Y, Z, ids.data_ptr(), offsets.data_ptr(),
                                           ids.numel(), pts.data_ptr(), n_pts, L.f3(anisotropy), baked.data_ptr(),
                                           L.ptr(dist), status.data_ptr(), L.stream_ptr(dev)))
    if int(status.item()) & L.STATUS_MISSING_ID:
        present = set(torch.unique(m).tolist()) - {0}
        missing = sorted(present - set(int(k) for k in skeletons))
        raise KeyError(missing[0] if missing else "mask id without a skeleton")
    if average:
        baked = average_baked_skeletons(baked.unsqueeze(0)).squeeze(0)
    return (baked, dist) if return_distance else baked


def skeleton_to_mask(skeletons: Dict[int, Tensor], shape: Tuple[int, int, int], device=None, radius: int = 7,
                     flank_radius: int = 3) -> Tensor:
    """Drop-in for skoots.lib.skeleton.skeleton_to_mask (:531-593): OR-stamps the 3-slice disk around
    every skeleton point; returns (1,X,Y,Z) fp32 of {0,1} on the skeletons' (CUDA) device."""
    if -1 in skeletons:
        return torch.zeros(tuple(shape), device=device)
    if not skeletons:
        return torch.zeros(tuple(shape)).unsqueeze(0)
    first = next(iter(skeletons.values()))
    dev, staged = L.compute_device(first)
    if staged:
        return skeleton_to_mask({k: L.stage_in(v, dev) for k, v in skeletons.items()}, shape, device, radius, flank_radius).cpu()
    X, Y, Z = (int(s) for s in shape)
    out = torch.zeros((X, Y, Z), dtype=torch.float32, device=dev)
    pts = torch.cat([v.to(device=dev, dtype=torch.float32).reshape(-1, 3) for v in skeletons.values()], 0).contiguous()
    off = get_cached_disk_coords(dev, radius, flank_radius).T.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        L.check(L.load().skb_stamp_disks(pts.data_ptr(), pts.shape[0], off.data_ptr(), off.shape[0], X, Y, Z,
                                         out.data_ptr(), L.stream_ptr(dev)))
    return out.unsqueeze(0)
