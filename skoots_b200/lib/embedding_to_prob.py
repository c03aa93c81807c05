"""`skoots.lib.embedding_to_prob.baked_embed_to_prob` on B200 (+ the fused vector->probability op)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _lib as L
from ._util import as_floats

_FLOATS = (torch.float32, torch.float16, torch.bfloat16)


class _EmbedProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embedding: Tensor, baked: Tensor, sigma_vals, eps: float):
        dev = embedding.device
        E = embedding.float().contiguous()
        S = baked if baked.dtype in _FLOATS else baked.float()
        S = S.contiguous()
        B, C = E.shape[:2]
        inner = E[0, 0].numel()
        out = torch.empty((B, 1) + tuple(E.shape[2:]), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().skb_embed_prob_fwd(E.data_ptr(), S.data_ptr(), L.dtype_code(S), B, C, inner,
                                                L.f3(sigma_vals), float(eps), out.data_ptr(), L.stream_ptr(dev)))
        ctx.save_for_backward(E, S, out)
        ctx.sigma_vals, ctx.eps = sigma_vals, eps
        ctx.in_dtypes = (embedding.dtype, baked.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        E, S, out = ctx.saved_tensors
        dev = E.device
        go = grad_out.float().contiguous()
        need_e, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gE = torch.empty_like(E) if need_e else None
        gS = torch.empty_like(S) if need_s else None
        B, C = E.shape[:2]
        with torch.cuda.device(dev):
            L.check(L.load().skb_embed_prob_bwd(E.data_ptr(), S.data_ptr(), L.dtype_code(S), out.data_ptr(), go.data_ptr(),
                                                B, C, E[0, 0].numel(), L.f3(ctx.sigma_vals), float(ctx.eps),
                                                L.ptr(gE), L.ptr(gS), L.stream_ptr(dev)))
        if gE is not None:
            gE = gE.to(ctx.in_dtypes[0])
        if gS is not None:
            gS = gS.to(ctx.in_dtypes[1])
        return gE, gS, None, None


def baked_embed_to_prob(embedding: Tensor, baked_skeletons: Tensor, sigma: Tensor, eps: float = 1e-16) -> Tensor:
    """exp(sum_c (E_c - S_c)^2 / (-2 (sigma_c + eps)^2)); (B,C,...) -> (B,1,...) fp32, C = 2 or 3
    (embedding_to_prob.py:5-51).  Differentiable w.r.t. embedding and baked_skeletons."""
    dev, staged = L.compute_device(embedding, baked_skeletons)
    if staged:
        return baked_embed_to_prob(L.stage_in(embedding, dev), L.stage_in(baked_skeletons, dev), sigma, eps).cpu()
    if embedding.shape != baked_skeletons.shape:
        raise RuntimeError(f"embedding {tuple(embedding.shape)} and baked_skeletons {tuple(baked_skeletons.shape)} differ")
    C = embedding.shape[1]
    if C not in (2, 3) or embedding.ndim != C + 2:
        raise RuntimeError(f"expected (B,2,X,Y) or (B,3,X,Y,Z), got {tuple(embedding.shape)}")
    sigma_vals = as_floats(sigma, C)
    if torch.is_grad_enabled() and (embedding.requires_grad or baked_skeletons.requires_grad):
        return _EmbedProb.apply(embedding, baked_skeletons, sigma_vals, eps)
    return _EmbedProb.forward(_NoCtx(), embedding, baked_skeletons, sigma_vals, eps)


class _NoCtx:
    def save_for_backward(self, *a):
        pass


class _VecProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vector: Tensor, baked: Tensor, scale_vals, sigma_vals, eps: float):
        dev = vector.device
        v = vector if vector.dtype in _FLOATS else vector.float()
        v = v.contiguous()
        S = baked if baked.dtype in _FLOATS else baked.float()
        S = S.contiguous()
        B, C = v.shape[:2]
        dims = list(v.shape[2:]) + [1] * (3 - (v.ndim - 2))
        out = torch.empty((B, 1) + tuple(v.shape[2:]), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().skb_vec_prob(v.data_ptr(), L.dtype_code(v), S.data_ptr(), L.dtype_code(S), B, C, dims[0],
                                          dims[1], dims[2], L.f3(scale_vals), L.f3(sigma_vals), float(eps),
                                          out.data_ptr(), 0, 0, L.stream_ptr(dev)))
        ctx.save_for_backward(v, S)
        ctx.args = (scale_vals, sigma_vals, eps, dims, vector.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        v, S = ctx.saved_tensors
        scale_vals, sigma_vals, eps, dims, in_dtype = ctx.args
        dev = v.device
        go = grad_out.float().contiguous()
        gv = torch.empty_like(v)
        B, C = v.shape[:2]
        with torch.cuda.device(dev):
            L.check(L.load().skb_vec_prob(v.data_ptr(), L.dtype_code(v), S.data_ptr(), L.dtype_code(S), B, C, dims[0],
                                          dims[1], dims[2], L.f3(scale_vals), L.f3(sigma_vals), float(eps), 0,
                                          go.data_ptr(), gv.data_ptr(), L.stream_ptr(dev)))
        return gv.to(in_dtype), None, None, None, None


def vector_to_prob(scale: Tensor, vector: Tensor, baked_skeletons: Tensor, sigma: Tensor, eps: float = 1e-16) -> Tensor:
    """baked_embed_to_prob(vector_to_embedding(scale, vector), baked, sigma) in one kernel: the
    fp32 embedding is never written to HBM (train/engine.py:465-466).  Differentiable w.r.t. vector."""
    dev, staged = L.compute_device(vector, baked_skeletons)
    if staged:
        return vector_to_prob(scale, L.stage_in(vector, dev), L.stage_in(baked_skeletons, dev), sigma, eps).cpu()
    C = vector.shape[1]
    if C not in (2, 3) or vector.ndim != C + 2 or vector.shape != baked_skeletons.shape:
        raise RuntimeError("vector and baked_skeletons must both be (B,2,X,Y) or (B,3,X,Y,Z)")
    scale_vals, sigma_vals = as_floats(scale, C), as_floats(sigma, C)
    if torch.is_grad_enabled() and vector.requires_grad:
        return _VecProb.apply(vector, baked_skeletons, scale_vals, sigma_vals, eps)
    return _VecProb.forward(_NoCtx(), vector, baked_skeletons, scale_vals, sigma_vals, eps)
