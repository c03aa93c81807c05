"""Synthetic "analytic tubes" volumes used by bench.py and the tests.

The reference ships no data; its hot path consumes (a) a u8 skeleton mask, (b) an fp16
3-vector field pointing at the object's skeleton and (c) for training, an integer instance
mask plus a ``Dict[int, Tensor[M,3]]`` of skeleton points (skoots/lib/eval.py:102-103,
skoots/train/dataloader.py:111-115).  This module draws all of them from straight tubes:

  tube i = segment a_i -> b_i;  a ~ U(box), dir ~ N(0,I) with dz*0.2, len ~ U(20,60)
  instance mask   : voxels closer than ``radius`` to the segment (nearest tube wins)
  skeleton mask   : voxels closer than ``skel_radius``
  vectors         : (nearest point on the segment - voxel) / scale, clipped to [-1,1]
  skeleton dict   : rounded points along the segment, one per unit length

Everything is torch, so the same code fills a 128x128x32 CPU fixture or one rank's Z-slab
of a 2048x2048x512 volume directly in HBM (``z_range`` selects the slab, ``xy_range`` a box; tubes are drawn
for the whole volume from the seed so every rank sees the same objects).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch


@dataclass
class TubeVolume:
    mask: torch.Tensor          # (X,Y,Zl) int32 instance ids (0 = background)
    skeleton: torch.Tensor      # (X,Y,Zl) uint8 {0,1}
    vectors: torch.Tensor       # (3,X,Y,Zl) float16 in [-1,1]
    skeletons: Dict[int, torch.Tensor]  # id -> (M,3) float32 points (global coords)
    shape: Tuple[int, int, int]  # global (X,Y,Z)
    z_range: Tuple[int, int]     # slab [z0,z1) of the global volume held here


def draw_tubes(shape, n_tubes: int, seed: int = 0, flat: bool = False):
    """Endpoints (n,3) a,b of the tubes, float64 numpy.  ``flat`` => dz = 0 (2-D mode, C5)."""
    rng = np.random.default_rng(seed)
    dims = np.asarray(shape, dtype=np.float64)
    a = rng.uniform(0.0, 1.0, size=(n_tubes, 3)) * dims
    d = rng.normal(size=(n_tubes, 3))
    d[:, 2] *= 0.0 if flat else 0.2
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    length = rng.uniform(20.0, 60.0, size=(n_tubes, 1))
    b = a + length * d
    return a, b


def make_tube_volume(
    shape: Tuple[int, int, int],
    n_tubes: int,
    seed: int = 0,
    device: str | torch.device = "cpu",
    scale: Tuple[float, float, float] = (60.0, 60.0, 12.0),
    radius: float = 4.0,
    skel_radius: float = 1.5,
    z_range: Optional[Tuple[int, int]] = None,
    xy_range: Optional[Tuple[Tuple[int, int], Tuple[int, int]]] = None,
    flat: bool = False,
    want_mask: bool = True,
    want_skeleton_dict: bool = True,
) -> TubeVolume:
    X, Y, Z = shape
    z0, z1 = (0, Z) if z_range is None else z_range
    (x0, x1), (y0, y1) = ((0, X), (0, Y)) if xy_range is None else xy_range  # only this box of the volume is filled
    Zl, Xl, Yl = z1 - z0, x1 - x0, y1 - y0
    dev = torch.device(device)
    a_np, b_np = draw_tubes(shape, n_tubes, seed, flat)

    best = torch.full((Xl, Yl, Zl), float("inf"), dtype=torch.float32, device=dev)
    mask = torch.zeros((Xl, Yl, Zl), dtype=torch.int32, device=dev) if want_mask else None
    skel = torch.zeros((Xl, Yl, Zl), dtype=torch.uint8, device=dev)
    vec = torch.zeros((3, Xl, Yl, Zl), dtype=torch.float16, device=dev)
    sc = torch.tensor(scale, dtype=torch.float32, device=dev)

    pad = radius + 1.0
    for i in range(n_tubes):
        a, b = a_np[i], b_np[i]
        lo = np.floor(np.minimum(a, b) - pad).astype(np.int64)
        hi = np.ceil(np.maximum(a, b) + pad).astype(np.int64) + 1
        lo = np.maximum(lo, [x0, y0, z0])
        hi = np.minimum(hi, [x1, y1, z1])
        if np.any(hi <= lo):
            continue
        gx = torch.arange(lo[0], hi[0], device=dev, dtype=torch.float32)[:, None, None]
        gy = torch.arange(lo[1], hi[1], device=dev, dtype=torch.float32)[None, :, None]
        gz = torch.arange(lo[2], hi[2], device=dev, dtype=torch.float32)[None, None, :]
        ab = b - a
        denom = float(ab @ ab) + 1e-12
        t = ((gx - a[0]) * ab[0] + (gy - a[1]) * ab[1] + (gz - a[2]) * ab[2]) / denom
        t = t.clamp_(0.0, 1.0)
        px, py, pz = a[0] + t * ab[0], a[1] + t * ab[1], a[2] + t * ab[2]
        dx, dy, dz = px - gx, py - gy, pz - gz
        dist = torch.sqrt(dx * dx + dy * dy + dz * dz)

        sl = (slice(lo[0] - x0, hi[0] - x0), slice(lo[1] - y0, hi[1] - y0), slice(lo[2] - z0, hi[2] - z0))
        cur = best[sl]
        win = (dist < radius) & (dist < cur)
        best[sl] = torch.where(win, dist, cur)
        if mask is not None:
            mask[sl] = torch.where(win, torch.full_like(mask[sl], i + 1), mask[sl])
        skel[sl] = torch.where(dist < skel_radius, torch.ones_like(skel[sl]), skel[sl])
        for c, dd in enumerate((dx, dy, dz)):
            vsl = (c,) + sl
            vnew = (dd / sc[c]).clamp_(-1.0, 1.0).to(torch.float16)
            vec[vsl] = torch.where(win, vnew, vec[vsl])
    del best

    skeletons: Dict[int, torch.Tensor] = {}
    if want_skeleton_dict:
        for i in range(n_tubes):
            a, b = a_np[i], b_np[i]
            n_pts = int(np.floor(np.linalg.norm(b - a))) + 1
            t = np.linspace(0.0, 1.0, n_pts)[:, None]
            pts = np.round(a[None, :] + t * (b - a)[None, :]).astype(np.float32)
            skeletons[i + 1] = torch.from_numpy(pts).to(dev)

    return TubeVolume(mask=mask, skeleton=skel, vectors=vec, skeletons=skeletons,
                      shape=(X, Y, Z), z_range=(z0, z1))
