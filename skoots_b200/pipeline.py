"""The post-UNet half of `skoots.lib.eval.eval` (skoots/lib/eval.py:223-284) as one device-side call.

`assemble_instances` = efficient_flood_fill + the crop loop {vector_to_embedding(N) ; += origin ;
index_skeleton_by_embed ; write interior}, fused: the skeleton mask is labelled in sparse form and
every voxel's walk + label gather happens in a single kernel, with no embedding or dense label
volume ever materialised in HBM.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L
from .lib._util import as_floats
from .lib.flood_fill import SparseLabels, label_components

EVAL_CROP = (500, 500, 50)      # skoots/lib/eval.py:248
EVAL_OVERLAP = (50, 50, 5)      # skoots/lib/eval.py:249
EVAL_N = 10                     # skoots/lib/eval.py:272


def gather_instances(vectors: Tensor, scale, labels, N: int = 1, decay: float = 1.0,
                     crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0),
                     out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.int32) -> Tensor:
    """vectors (3,X,Y,Z) f16/bf16/f32; labels = SparseLabels or a dense (X,Y,Z) int16/int32 volume."""
    dev = L.require_cuda(vectors)
    if vectors.ndim != 4 or vectors.shape[0] != 3:
        raise RuntimeError(f"vectors must be (3,X,Y,Z), got {tuple(vectors.shape)}")
    if vectors.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        vectors = vectors.float()
    vectors = vectors.contiguous()
    _, X, Y, Z = vectors.shape
    crop = (X, Y, Z) if crop is None else tuple(int(c) for c in crop)
    overlap = tuple(int(o) for o in overlap)
    if out is None:
        out = torch.empty((X, Y, Z), dtype=out_dtype, device=dev)
    assert out.is_contiguous() and tuple(out.shape) == (X, Y, Z) and out.dtype in (torch.int32, torch.int16)
    ws_ptr, dense_ptr, dense_code = 0, 0, 0
    if isinstance(labels, SparseLabels):
        assert labels.shape == (X, Y, Z), "label workspace was built for another volume"
        ws_ptr = labels.workspace.data_ptr()
    else:
        L.require_cuda(labels)
        dense = labels.reshape(X, Y, Z).contiguous()
        if dense.dtype not in (torch.int16, torch.int32, torch.uint8):
            dense = dense.to(torch.int32)
        dense_ptr, dense_code = dense.data_ptr(), L.dtype_code(dense)
    with torch.cuda.device(dev):
        L.check(L.load().skb_assemble(vectors.data_ptr(), L.dtype_code(vectors), X, Y, Z, L.f3(as_floats(scale, 3)),
                                      int(N), float(decay), L.i3(crop), L.i3(overlap), ws_ptr, dense_ptr, dense_code,
                                      out.data_ptr(), L.dtype_code(out), L.stream_ptr(dev)))
    return out


def assemble_instances(skeleton_mask: Tensor, vectors: Tensor, scale, N: int = 1, decay: float = 1.0,
                       crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0),
                       out_dtype: torch.dtype = torch.int32, check: bool = True,
                       workspace: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """skeleton_mask (X,Y,Z) or (1,X,Y,Z) u8/bool/int16; vectors (3,X,Y,Z) -> instance labels (X,Y,Z).

    crop=None treats the whole volume as one crop (the lib functions applied directly);
    crop=EVAL_CROP, overlap=EVAL_OVERLAP, N=EVAL_N, out_dtype=int16 reproduces eval()."""
    mask = skeleton_mask.squeeze(0) if skeleton_mask.ndim == 4 else skeleton_mask
    sparse = label_components(mask, planar=False, label_base=2, workspace=workspace, check=check)
    return gather_instances(vectors, scale, sparse, N=N, decay=decay, crop=crop, overlap=overlap, out=out,
                            out_dtype=out_dtype)


class HostAssembler:
    """End-to-end form of `assemble_instances` for volumes that live in HOST memory (the reference
    keeps them in zarr / numpy, skoots/lib/eval.py:102-103,223,245): copies the u8 skeleton mask
    and the fp16 vectors host->device, labels + gathers on the GPU, and copies the instance mask
    back.  Device buffers and the CCL workspace are allocated once and reused across calls.

    Uploads run on a copy stream: the CCL starts as soon as the mask has landed and overlaps the
    (6x larger) vector upload; the gather waits for the vectors.
    """

    def __init__(self, shape: Tuple[int, int, int], device="cuda:0", vec_dtype=torch.float16,
                 out_dtype=torch.int32):
        X, Y, Z = shape
        self.shape, self.dev = (X, Y, Z), torch.device(device)
        self.mask = torch.empty((X, Y, Z), dtype=torch.uint8, device=self.dev)
        self.vec = torch.empty((3, X, Y, Z), dtype=vec_dtype, device=self.dev)
        self.out = torch.empty((X, Y, Z), dtype=out_dtype, device=self.dev)
        self.workspace = None
        self.copy_stream = torch.cuda.Stream(self.dev)

    def __call__(self, mask_host: Tensor, vec_host: Tensor, scale, out_host: Tensor, N: int = 1, decay: float = 1.0,
                 crop=None, overlap=(0, 0, 0)) -> Tensor:
        X, Y, Z = self.shape
        main = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            self.mask.copy_(mask_host.reshape(X, Y, Z), non_blocking=True)
            mask_ready = torch.cuda.Event()
            mask_ready.record(cs)
            self.vec.copy_(vec_host.reshape(3, X, Y, Z), non_blocking=True)
            vec_ready = torch.cuda.Event()
            vec_ready.record(cs)
        main.wait_event(mask_ready)
        sparse = label_components(self.mask, label_base=2, workspace=self.workspace, check=False)
        self.workspace = sparse.workspace
        main.wait_event(vec_ready)
        gather_instances(self.vec, scale, sparse, N=N, decay=decay, crop=crop, overlap=overlap, out=self.out)
        out_host.reshape(X, Y, Z).copy_(self.out, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        sparse.check()
        return out_host


def tile_epilogue(unet_out: Tensor, vectors: Tensor, skeleton: Tensor, origin: Sequence[int],
                  overlap: Sequence[int] = (50, 50, 5), threshold: float = 0.8) -> None:
    """skoots/lib/eval.py:145-176 in one kernel: unet_out (1,C>=5,x,y,z) -> masked fp16 vectors and the
    dilated, thresholded u8 skeleton, written into the interior of the tile's slot of the device-
    resident whole-volume arrays `vectors` (3,X,Y,Z) fp16 and `skeleton` (1,X,Y,Z) or (X,Y,Z) uint8."""
    dev = L.require_cuda(unet_out, vectors, skeleton)
    if unet_out.ndim != 5 or unet_out.shape[0] != 1 or unet_out.shape[1] < 5:
        raise RuntimeError(f"unet_out must be (1,C>=5,x,y,z), got {tuple(unet_out.shape)}")
    if unet_out.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        unet_out = unet_out.float()
    unet_out = unet_out.contiguous()
    assert vectors.dtype == torch.float16 and vectors.is_contiguous() and vectors.shape[0] == 3
    assert skeleton.dtype == torch.uint8 and skeleton.is_contiguous()
    X, Y, Z = vectors.shape[1:]
    assert skeleton.numel() == X * Y * Z
    _, C, tx, ty, tz = unet_out.shape
    with torch.cuda.device(dev):
        L.check(L.load().skb_tile_epilogue(unet_out.data_ptr(), L.dtype_code(unet_out), C, L.i3((tx, ty, tz)),
                                           L.i3(origin), L.i3(overlap), float(threshold), vectors.data_ptr(),
                                           skeleton.data_ptr(), X, Y, Z, L.stream_ptr(dev)))
