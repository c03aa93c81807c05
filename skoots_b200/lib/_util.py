from __future__ import annotations

from typing import List, Sequence

import torch

_SCALE_CACHE: dict = {}


def as_floats(values, n: int | None = None) -> List[float]:
    """`scale`/`sigma`-style arguments arrive as tensors (CPU or CUDA, any dtype), lists or tuples.
    The reference converts with `.float()` (vector_to_embedding.py:90); a CUDA tensor costs one
    small D2H copy, cached on (storage, version) so a training loop pays it once."""
    if isinstance(values, torch.Tensor):
        key = (values.data_ptr(), values._version, values.dtype, tuple(values.shape), str(values.device))
        hit = _SCALE_CACHE.get(key)
        if hit is None:
            hit = [float(v) for v in values.detach().float().reshape(-1).tolist()]
            if len(_SCALE_CACHE) > 64:
                _SCALE_CACHE.clear()
            _SCALE_CACHE[key] = hit
        out = list(hit)
    else:
        out = [float(torch.tensor(v, dtype=torch.float32)) for v in values]
    if n is not None and len(out) != n:
        raise ValueError(f"expected {n} values, got {len(out)}")
    return out
