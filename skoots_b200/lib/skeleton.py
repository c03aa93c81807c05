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


def _pack_batch(skeletons_list, device):
    """id / offset tables of a batch of skeleton dicts: built on the host from the dicts' keys and the tensors' SHAPES
    (no device read), sent up as ONE small int32 array; the points of all samples are concatenated on the device."""
    ids, lens, begin, tensors = [], [], [0], []
    for sk in skeletons_list:
        keys = sorted(int(k) for k in sk.keys())
        for k in keys:
            t = sk[k]
            ids.append(k)
            lens.append(int(t.shape[0]))
            if t.shape[0]:
                tensors.append(t)
        begin.append(len(ids))
    n_ids, n_pts = len(ids), int(sum(lens))
    table = np.zeros(n_ids + len(begin) + n_ids + 1, dtype=np.int32)
    table[:n_ids] = ids
    table[n_ids:n_ids + len(begin)] = begin
    if n_ids:
        np.cumsum(lens, out=table[n_ids + len(begin) + 1:])
    table_d = torch.from_numpy(table).to(device, non_blocking=True)
    pts = torch.zeros((max(n_pts, 1), 4), dtype=torch.float32, device=device)
    if n_pts:
        ready = all(t.device == device and t.dtype == torch.float32 and t.ndim == 2 for t in tensors)  # the usual case: no per-tensor work
        pts[:n_pts, :3].copy_(torch.cat(tensors if ready else
                                        [t.to(device=device, dtype=torch.float32).reshape(-1, 3) for t in tensors], 0))
    return table_d[:n_ids], table_d[n_ids:n_ids + len(begin)], table_d[n_ids + len(begin):], pts, n_ids, n_pts


def _next_power_of_2(x: int) -> int:
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


def bake_skeletons_batch(masks, skeletons_list, anisotropy: Tuple[float, float, float] = (1.0, 1.0, 1.0), average: bool = True,
                         return_distance: bool = False, check: bool = True, triton_compat: bool = False):
    """`bake_skeleton` for a whole batch in ONE launch: masks (B,X,Y,Z) tensor (or a list of (X,Y,Z) / (1,X,Y,Z) tensors of
    one shape), skeletons_list = one `Dict[int, Tensor[M,3]]` per sample.  Returns (B,3,X,Y,Z) fp32 (and (B,1,X,Y,Z)
    distances).  The status word is read ONCE for the batch (check=False: not at all; the caller owns the check).
    Semantics per sample = the reference's CPU path (skeleton.py:370-445) followed by average_baked_skeletons;
    triton_compat=True = the reference's Triton kernel instead (skeleton.py:51-367, what it runs for CUDA masks): values
    are fp16-representable, a missing id gives zeros and raises nothing."""
    if not isinstance(masks, torch.Tensor):
        masks = torch.stack([m.squeeze(0) if m.ndim == 4 else m for m in masks])
    dev = L.require_cuda(masks)
    assert masks.ndim == 4 and masks.shape[0] == len(skeletons_list), "one skeleton dict per sample"
    m = masks if masks.dtype in (torch.int32, torch.int16, torch.uint8) else masks.to(torch.int32)
    m = m.contiguous()
    B, X, Y, Z = m.shape
    ids, begin, offsets, pts, n_ids, n_pts = _pack_batch(skeletons_list, dev)
    blocks = None
    if triton_compat:  # SKEL_BLOCK_SIZE of the reference's launch per sample (skeleton.py:302,361); 0 = it returns zeros (:304)
        longest = [max((int(v.shape[0]) for v in sk.values()), default=0) for sk in skeletons_list]
        blocks = torch.tensor([_next_power_of_2(n) if n else 0 for n in longest], dtype=torch.int32).to(dev, non_blocking=True)
    baked = torch.empty((B, 3, X, Y, Z), dtype=torch.float32, device=dev)
    dist = torch.empty((B, 1, X, Y, Z), dtype=torch.float32, device=dev) if return_distance else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().skb_bake_skeletons(m.data_ptr(), L.dtype_code(m), B, X, Y, Z, ids.data_ptr(), begin.data_ptr(),
                                            offsets.data_ptr(), n_ids, pts.data_ptr(), n_pts, L.f3(anisotropy), int(bool(average)),
                                            L.ptr(blocks), baked.data_ptr(), L.ptr(dist), status.data_ptr(), L.stream_ptr(dev)))
    if check and not triton_compat and int(status.item()) & L.STATUS_MISSING_ID:
        for b in range(B):
            present = set(torch.unique(m[b]).tolist()) - {0}
            missing = sorted(present - set(int(k) for k in skeletons_list[b]))
            if missing:
                raise KeyError(missing[0])
        raise KeyError("mask id without a skeleton")
    return (baked, dist) if return_distance else baked


def bake_skeleton(masks: Tensor, skeletons: Dict[int, Tensor], anisotropy: Tuple[float, float, float] = (1.0, 1.0, 1.0),
                  average: bool = True, device: str = "cpu", return_distance: bool = False, triton_compat=False):
    """Drop-in for skoots.lib.skeleton.bake_skeleton (:448-528) with the CPU/torch semantics
    (anisotropy scales coordinates, first minimum wins, fp32 out — SURVEY A.5).  `device` is
    accepted and ignored like the reference's positional mix-up (:507); the work runs on
    masks.device (host tensors are staged through the GPU).  return_distance=True also returns the (1,X,Y,Z)
    distance.  One kernel launch (nearest point + the masked 27-mean fused) and one status read — the reference raises
    KeyError synchronously for a mask id without a skeleton (:422), so does this; `bake_skeletons_batch` does a whole
    batch with one launch and one read.

    triton_compat: the reference answers a CUDA mask with its Triton kernel (:505-512), whose results differ from its CPU
    path (see skb_train.cu, bake_nearest_triton).  True = those semantics (fp16 baked when average=False, fp16 distance,
    no KeyError); "auto" = the reference's own dispatch (Triton semantics for CUDA masks, CPU semantics for host
    tensors) — what `patch_skoots(bug_compatible=True)` binds; False (default) = the CPU semantics everywhere."""
    dev, staged = L.compute_device(masks)
    if triton_compat == "auto":
        triton_compat = not staged
    if staged:
        return L.stage_out(bake_skeleton(L.stage_in(masks, dev), skeletons, anisotropy, average, device, return_distance,
                                         triton_compat), True)
    if -1 in skeletons:
        x, y, z = masks.shape[-3:]
        return torch.zeros((3, x, y, z), device=dev, dtype=torch.float16)
    if masks.ndim == 4 and masks.shape[0] == 1:
        masks = masks.squeeze(0)
    assert masks.ndim == 3, f"masks must be 3d with no batch. not {masks.shape=}"
    out = bake_skeletons_batch(masks.unsqueeze(0), [skeletons], anisotropy, average, return_distance, triton_compat=bool(triton_compat))
    baked, dist = (out[0][0], out[1][0]) if return_distance else (out[0], None)
    if triton_compat:  # the Triton launcher's dtypes (:287-293): fp16, widened only by the averaging (:519-523)
        baked = baked if average else baked.to(torch.float16)
        dist = None if dist is None else dist.to(torch.float16)
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
    vals = list(skeletons.values())
    ready = all(v.device == dev and v.dtype == torch.float32 and v.ndim == 2 for v in vals)  # the usual case: one cat, no per-tensor work
    pts = torch.cat(vals if ready else [v.to(device=dev, dtype=torch.float32).reshape(-1, 3) for v in vals], 0).contiguous()
    key = ("i32", str(dev), int(radius), int(flank_radius))
    off = _DISK_CACHE.get(key)
    if off is None:  # (S,3) int32 form of the cached stamp, built once per (device, radius, flank)
        off = _DISK_CACHE[key] = get_cached_disk_coords(dev, radius, flank_radius).T.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        L.check(L.load().skb_stamp_disks(pts.data_ptr(), pts.shape[0], off.data_ptr(), off.shape[0], X, Y, Z,
                                         out.data_ptr(), L.stream_ptr(dev)))
    return out.unsqueeze(0)
