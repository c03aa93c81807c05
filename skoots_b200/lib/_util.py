from __future__ import annotations

import weakref
from typing import List

import torch

_CACHE: dict = {}


def as_floats(values, n: int | None = None) -> List[float]:
    """`scale`/`sigma`-style arguments arrive as tensors (CPU or CUDA, any dtype), lists or tuples.
    The reference converts with `.float()` (vector_to_embedding.py:90).  Reading a CUDA tensor costs
    one small D2H copy; it is cached per tensor object (weak reference + in-place version counter)
    so a training loop that passes the same `vector_scale` every step pays it once."""
    if isinstance(values, torch.Tensor):
        if values.is_cuda:
            key = id(values)
            hit = _CACHE.get(key)
            if hit is not None and hit[0]() is values and hit[1] == values._version:
                out = list(hit[2])
            else:
                out = [float(v) for v in values.detach().float().reshape(-1).tolist()]
                if len(_CACHE) > 256:
                    _CACHE.clear()
                _CACHE[key] = (weakref.ref(values), values._version, tuple(out))
        else:
            out = [float(v) for v in values.detach().float().reshape(-1).tolist()]
    else:
        out = [float(torch.tensor(v, dtype=torch.float32)) for v in values]
    if n is not None and len(out) != n:
        raise ValueError(f"expected {n} values, got {len(out)}")
    return out
