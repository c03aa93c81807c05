"""`skoots.lib.morphology` on B200 (reference: skoots/lib/morphology.py:130-199)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _lib as L

_MAX333, _MAX331, _MIN333 = 0, 1, 2


def _stencil(image: Tensor, op: int) -> Tensor:
    dev, staged = L.compute_device(image)
    if staged:
        return _stencil(L.stage_in(image, dev), op).cpu()
    if image.ndim != 5:
        raise RuntimeError(f"expected a 5D (B,C,X,Y,Z) tensor, got {tuple(image.shape)}")
    src = image.float().contiguous()
    b, c, X, Y, Z = src.shape
    out = torch.empty_like(src)
    with torch.cuda.device(dev):
        L.check(L.load().skb_stencil3(src.data_ptr(), out.data_ptr(), b * c, X, Y, Z, op, L.stream_ptr(dev)))
    return out


def binary_dilation(image: Tensor) -> Tensor:
    """3x3x3 zero-padded max over (B,C,X,Y,Z) (morphology.py:155-175)."""
    return _stencil(image, _MAX333)


def binary_dilation_2d(image: Tensor) -> Tensor:
    """3x3x1 zero-padded max (morphology.py:178-199)."""
    return _stencil(image, _MAX331)


def binary_erosion(image: Tensor) -> Tensor:
    """3x3x3 zero-padded min; like the reference it returns shape (1, B*C, X, Y, Z) (:152)."""
    b, c, X, Y, Z = image.shape
    return _stencil(image, _MIN333).reshape(1, b * c, X, Y, Z)
