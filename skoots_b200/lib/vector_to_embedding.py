"""`skoots.lib.vector_to_embedding` on B200 (reference: skoots/lib/vector_to_embedding.py)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _lib as L
from ._util import as_floats


class _Vec2Embed(torch.autograd.Function):
    """phi = idx + v*s (N=1).  d phi_c / d v_c = s_c  (vector_to_embedding.py:104-105)."""

    @staticmethod
    def forward(ctx, vector: Tensor, scale_vals, ndim_spatial: int):
        ctx.scale_vals = scale_vals
        ctx.vec_dtype = vector.dtype
        return _forward(vector, scale_vals, 1, 1.0)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        grad_out = grad_out.contiguous().float()
        B, C = grad_out.shape[:2]
        inner = grad_out[0, 0].numel()
        grad_vec = torch.empty(grad_out.shape, dtype=ctx.vec_dtype, device=grad_out.device)
        with torch.cuda.device(grad_out.device):  # the launch must go to the tensor's GPU, not the current one
            L.check(L.load().skb_vec_embed_bwd(grad_out.data_ptr(), B, C, inner, L.f3(ctx.scale_vals),
                                               grad_vec.data_ptr(), L.dtype_code(grad_vec),
                                               L.stream_ptr(grad_out.device)))
        return grad_vec, None, None


def _forward(vector: Tensor, scale_vals, N: int, decay: float) -> Tensor:
    dev = L.require_cuda(vector)
    if vector.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        vector = vector.float()
    vector = vector.contiguous()
    lib = L.load()
    out = torch.empty(vector.shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if vector.ndim == 5:
            B, C, X, Y, Z = vector.shape
            assert C == 3, f"3-D vector field must have 3 channels, not {C}"
            L.check(lib.skb_vec_embed3d(vector.data_ptr(), L.dtype_code(vector), B, X, Y, Z, L.f3(scale_vals),
                                        int(N), float(decay), out.data_ptr(), L.stream_ptr(dev)))
        else:
            B, C, X, Y = vector.shape
            assert C == 2, f"2-D vector field must have 2 channels, not {C}"
            L.check(lib.skb_vec_embed2d(vector.data_ptr(), L.dtype_code(vector), B, X, Y, L.f3(scale_vals),
                                        out.data_ptr(), L.stream_ptr(dev)))
    return out


def vector_to_embedding(scale: Tensor, vector: Tensor, N: int = 1, decay: float = 1.0) -> Tensor:
    """Same contract as skoots.lib.vector_to_embedding.vector_to_embedding (:135-174):
    (B,3,X,Y,Z) -> fp32 embedding with N-1 crop-local hops, or (B,2,X,Y) with N == 1.
    A HOST `vector` (eval() passes CPU crops, skoots/lib/eval.py:271) is staged through the GPU and the
    embedding comes back as a host tensor."""
    dev, staged = L.compute_device(vector)
    if staged:
        return vector_to_embedding(scale, L.stage_in(vector, dev), N, decay).cpu()
    if vector.ndim == 4:
        assert decay == 1.0, f'decay parameter only valid for 5D tensor'
        assert N == 1, f'N must be equal to 1 for 4D tensors.'
    elif vector.ndim != 5:
        raise RuntimeError(f"vector must be a 4D or 5D tensor, not {vector.ndim}D")
    scale_vals = as_floats(scale, vector.ndim - 2)
    if vector.requires_grad and torch.is_grad_enabled():
        if N != 1:
            raise NotImplementedError("autograd through vector_to_embedding is implemented for N == 1 (the training path)")
        return _Vec2Embed.apply(vector, scale_vals, vector.ndim - 2)
    return _forward(vector, scale_vals, N, decay)
