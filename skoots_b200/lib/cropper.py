"""`skoots.lib.cropper` host-side mirror (reference: skoots/lib/cropper.py).

Pure index arithmetic — no device work.  Keeps the reference's observable quirks: a crop larger than
the image is clamped and the CALLER'S `crop_size` list is mutated (cropper.py:13-16,81-84; eval()
relies on that at eval.py:129,162-164); the last crop of an axis is shifted back to dim-size and may
be yielded more than once.  One deliberate difference: where the reference loops forever
(crop - 2*overlap < 0, SURVEY B#16) this raises ValueError.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor


def _clamp_in_place(image_shape, crop_size: List[int]) -> None:
    for i in range(len(crop_size)):
        crop_size[i] = crop_size[i] if crop_size[i] < image_shape[i + 1] else image_shape[i + 1]


def _check(image_shape, crop_size, overlap) -> None:
    assert len(image_shape) - 1 == len(crop_size) == len(overlap) == 3, (
        f"Image Shape must equal the shape of the crop.\n{image_shape=}, {crop_size=}"
        f"{overlap=}"
    )
    for c, o, d in zip(crop_size, overlap, ("x", "y", "z")):
        assert c - o * 2 != 0, f"Overlap in {d} dimmension cannot be equal to or larger than crop size... {o*2=} < {c}"
        if c - o * 2 < 0:
            raise ValueError(f"crop size {c} is smaller than twice the overlap {o} in {d}: the reference would never terminate")


def _origins(dim: int, size: int, overlap: int) -> List[int]:
    out, o = [], 0
    while o < dim:
        out.append(o if o + size <= dim else dim - size)
        o += size - 2 * overlap
    return out


def get_total_num_crops(image_shape, crop_size: List[int], overlap: Optional[Tuple[int]]) -> int:
    """number of crops `crops` will yield (cropper.py:8-55)."""
    _clamp_in_place(image_shape, crop_size)
    _check(image_shape, crop_size, overlap)
    total = 1
    for a in range(3):
        total *= len(_origins(image_shape[a + 1], crop_size[a], overlap[a]))
    return total


def crops(image: Tensor, crop_size: List[int], overlap: Optional[Sequence[int]] = (0, 0, 0), device="cpu"
          ) -> Iterator[Tuple[Tensor, List[int]]]:
    """yields (crop (1,C,x,y,z) on `device`, [x,y,z] origin); x outermost, z innermost (cropper.py:58-144)."""
    shape = image.shape
    _clamp_in_place(shape, crop_size)
    _check(shape, crop_size, overlap)
    for x in _origins(shape[1], crop_size[0], overlap[0]):
        for y in _origins(shape[2], crop_size[1], overlap[1]):
            for z in _origins(shape[3], crop_size[2], overlap[2]):
                piece = image[:, x:x + crop_size[0], y:y + crop_size[1], z:z + crop_size[2]]
                piece = torch.from_numpy(piece) if isinstance(piece, np.ndarray) else piece
                yield piece.unsqueeze(0).to(device, non_blocking=True), [x, y, z]
