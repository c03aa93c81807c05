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
from .lib.flood_fill import SparseLabels, _as_mask, label_components, launch_label, new_sparse

EVAL_CROP = (500, 500, 50)      # skoots/lib/eval.py:248
EVAL_OVERLAP = (50, 50, 5)      # skoots/lib/eval.py:249
EVAL_N = 10                     # skoots/lib/eval.py:272


def gather_instances(vectors: Tensor, scale, labels, N: int = 1, decay: float = 1.0,
                     crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0),
                     out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.int32,
                     voxel_range: Optional[Tuple[int, int]] = None) -> Tensor:
    """vectors (3,X,Y,Z) f16/bf16/f32; labels = SparseLabels or a dense (X,Y,Z) int16/int32 volume.
    voxel_range=(first, count) restricts the pass to that stretch of the flat (X,Y,Z) index."""
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
    first, count = (0, X * Y * Z) if voxel_range is None else (int(voxel_range[0]), int(voxel_range[1]))
    with torch.cuda.device(dev):
        L.check(L.load().skb_assemble_range(vectors.data_ptr(), L.dtype_code(vectors), X, Y, Z, L.f3(as_floats(scale, 3)),
                                            int(N), float(decay), L.i3(crop), L.i3(overlap), ws_ptr, dense_ptr, dense_code,
                                            out.data_ptr(), L.dtype_code(out), first, count, L.stream_ptr(dev)))
    return out


def gather_instances_2d(vectors: Tensor, scale, labels, out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.int32) -> Tensor:
    """2-D mode (BASELINE configs[4]): vectors (S,2,X,Y) f16/bf16/f32 of S independent images; labels = planar
    SparseLabels of the (S,X,Y) stack or a dense (S,X,Y) label stack -> (S,X,Y) labels.  The reference's 2-D gather is
    `index_skeleton_by_embed` per slice with Z = 1 on `_vec2embed2D`'s embedding (skeleton.py:656-695,
    vector_to_embedding.py:50-76)."""
    dev = L.require_cuda(vectors)
    if vectors.ndim != 4 or vectors.shape[1] != 2:
        raise RuntimeError(f"vectors must be (S,2,X,Y), got {tuple(vectors.shape)}")
    if vectors.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        vectors = vectors.float()
    vectors = vectors.contiguous()
    S, _, X, Y = vectors.shape
    if out is None:
        out = torch.empty((S, X, Y), dtype=out_dtype, device=dev)
    assert out.is_contiguous() and tuple(out.shape) == (S, X, Y) and out.dtype in (torch.int32, torch.int16)
    ws_ptr, dense_ptr, dense_code = 0, 0, 0
    if isinstance(labels, SparseLabels):
        assert labels.shape == (S, X, Y), "label workspace was built for another stack"
        ws_ptr = labels.workspace.data_ptr()
    else:
        L.require_cuda(labels)
        dense = labels.reshape(S, X, Y).contiguous()
        if dense.dtype not in (torch.int16, torch.int32, torch.uint8):
            dense = dense.to(torch.int32)
        dense_ptr, dense_code = dense.data_ptr(), L.dtype_code(dense)
    with torch.cuda.device(dev):
        L.check(L.load().skb_assemble_planar(vectors.data_ptr(), L.dtype_code(vectors), S, X, Y, L.f3(as_floats(scale, 2)), ws_ptr,
                                             dense_ptr, dense_code, out.data_ptr(), L.dtype_code(out), L.stream_ptr(dev)))
    return out


def assemble_instances_2d(skeleton_masks: Tensor, vectors: Tensor, scale, out_dtype: torch.dtype = torch.int32,
                          workspace: Optional[Tensor] = None, out: Optional[Tensor] = None, check: bool = True,
                          label_base: int = 0) -> Tensor:
    """2-D mode end to end: per-slice connected components (4-connectivity, numbering restarting per slice at
    label_base + 1 — `scipy.ndimage.label` on each plane, utils/flood_and_stitch.py:63-69) of the (S,X,Y) u8 stack,
    then the fused 2-D gather.  Returns (S,X,Y) labels."""
    sparse = label_components(skeleton_masks, planar=True, label_base=label_base, workspace=workspace, check=check)
    return gather_instances_2d(vectors, scale, sparse, out=out, out_dtype=out_dtype)


_CHAIN_STREAMS = {}


def chain_stream(dev: torch.device) -> torch.cuda.Stream:
    """the high-priority stream the labelling chain runs on while the gather's stream phase owns the rest of the GPU."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _CHAIN_STREAMS.get(key)
    if st is None:
        st = _CHAIN_STREAMS[key] = torch.cuda.Stream(dev, priority=-1)
    return st


def split_eligible(shape, vectors: Tensor, N: int, crop, overlap) -> bool:
    """the stream/resolve split covers the headline mode: N = 1, the whole volume as one crop, Z a multiple
    of 64 and the volume a multiple of 256 voxels (include/skoots_b200.h: skb_assemble_stream)."""
    X, Y, Z = shape
    whole = crop is None or all(int(c) >= d for c, d in zip(crop, shape))
    return (N == 1 and whole and not any(int(o) for o in overlap) and Z % 64 == 0 and (X * Y * Z) % 256 == 0
            and vectors.dtype in (torch.float16, torch.bfloat16, torch.float32) and vectors.is_contiguous()
            and vectors.data_ptr() % 16 == 0 and (X * Y * Z * vectors.element_size()) % 16 == 0)


def stream_ctas_default() -> int:
    """CTAs per SM of the stream phase when it is run as a persistent grid; 0 (default) = one CTA per 16 chunks.
    Measured on B200 (profiles/README.md): a persistent stream of 2-3 CTAs per SM cannot saturate HBM on its own
    (3.9 ms vs 3.2 ms at 2048x2048x512) and does not make the overlap with the labelling chain pay either."""
    import os
    return int(os.environ.get("SKB_STREAM_CTAS", "0"))


def assemble_split(mask: Tensor, vectors: Tensor, scale, sparse: SparseLabels, out: Tensor,
                   group_flags: Optional[Tensor] = None, timers=None, trace: Optional[dict] = None,
                   stream_ctas: Optional[int] = None) -> Tensor:
    """One N = 1 whole-volume pass with the gather split around the labelling:

        current stream : pack mask->bits | stream phase (6 B/voxel in, zeros out, flags work groups) | resolve
        chain stream   :                 | tile union-find, boundary unions, numbering (latency-bound) |

    Writes exactly what label_components + gather_instances write.  timers = (e0, e1): CUDA events recorded
    around the stream phase on the current stream."""
    dev = vectors.device
    X, Y, Z = sparse.shape
    lib = L.load()
    if group_flags is None:
        group_flags = torch.empty(X * Y * Z // 256, dtype=torch.int32, device=dev)
    main, side = torch.cuda.current_stream(dev), chain_stream(dev)

    def mark(name, stream):  # trace: name -> timing event (profiles/split_timeline.py)
        if trace is not None:
            trace[name] = torch.cuda.Event(enable_timing=True)
            trace[name].record(stream)

    mark("begin", main)
    launch_label(mask, sparse, False, 2, L.CCL_PHASE_PACK)
    bits_ready = torch.cuda.Event()
    bits_ready.record(main)
    mark("packed", main)
    side.wait_event(bits_ready)
    with torch.cuda.stream(side):
        mark("chain_begin", side)
        launch_label(mask, sparse, False, 2, L.CCL_PHASE_LABEL)
        labelled = torch.cuda.Event()
        labelled.record(side)
        mark("chain_end", side)
    with torch.cuda.device(dev):
        if timers is not None:
            timers[0].record(main)
        L.check(lib.skb_assemble_stream(vectors.data_ptr(), L.dtype_code(vectors), X, Y, Z, 0, Z, sparse.workspace.data_ptr(),
                                        group_flags.data_ptr(), out.data_ptr(), L.dtype_code(out),
                                        stream_ctas_default() if stream_ctas is None else int(stream_ctas), main.cuda_stream))
        if timers is not None:
            timers[1].record(main)
        mark("streamed", main)
        main.wait_event(labelled)
        mark("joined", main)
        L.check(lib.skb_assemble_resolve(vectors.data_ptr(), L.dtype_code(vectors), X, Y, Z, 0, Z, L.f3(as_floats(scale, 3)),
                                         sparse.workspace.data_ptr(), 0, 0, group_flags.data_ptr(), out.data_ptr(),
                                         L.dtype_code(out), main.cuda_stream))
    mark("resolved", main)
    return out


def assemble_instances(skeleton_mask: Tensor, vectors: Tensor, scale, N: int = 1, decay: float = 1.0,
                       crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0),
                       out_dtype: torch.dtype = torch.int32, check: bool = True,
                       workspace: Optional[Tensor] = None, out: Optional[Tensor] = None,
                       fused: Optional[bool] = None) -> Tensor:
    """skeleton_mask (X,Y,Z) or (1,X,Y,Z) u8/bool/int16; vectors (3,X,Y,Z) -> instance labels (X,Y,Z).

    crop=None treats the whole volume as one crop (the lib functions applied directly);
    crop=EVAL_CROP, overlap=EVAL_OVERLAP, N=EVAL_N, out_dtype=int16 reproduces eval().
    fused=None / True: label, then one fused gather (the fastest form measured);
    fused=False: the stream/resolve split with the gather's stream phase overlapped with the labelling
    (kept as a measured alternative: bit-identical, slower on B200 because the labelling kernels fill the SMs)."""
    mask = skeleton_mask.squeeze(0) if skeleton_mask.ndim == 4 else skeleton_mask
    if not mask.is_cuda and not vectors.is_cuda:
        # HOST volumes (what eval() holds, skoots/lib/eval.py:102-103,223): the pipelined upload / label / gather / download
        dev, _ = L.compute_device(mask, vectors)
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else (mask if mask.dtype == torch.uint8 else mask.gt(0).view(torch.uint8))
        v = vectors if vectors.dtype in (torch.float16, torch.bfloat16, torch.float32) else vectors.float()
        want = out.dtype if out is not None else out_dtype
        host_out = out if out is not None else torch.empty(tuple(m.shape), dtype=want)
        runner = HostAssembler(tuple(m.shape), dev, vec_dtype=v.dtype, out_dtype=want)
        return runner(m.contiguous(), v.contiguous(), scale, host_out, N=N, decay=decay, crop=crop, overlap=overlap, check=check)
    if fused is False:  # the opt-in split; the default path below stays as lean as it was (small volumes are launch-bound)
        dev = L.require_cuda(mask, vectors)
        if vectors.ndim != 4 or vectors.shape[0] != 3:
            raise RuntimeError(f"vectors must be (3,X,Y,Z), got {tuple(vectors.shape)}")
        if vectors.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            vectors = vectors.float()
        vectors = vectors.contiguous()
        shape = tuple(vectors.shape[1:])
        if not split_eligible(shape, vectors, N, crop, overlap):
            raise L.SkootsB200Error("the stream/resolve split needs N = 1, one whole-volume crop, Z % 64 == 0 and V % 256 == 0")
        mask = _as_mask(mask)
        sparse = new_sparse(shape, dev, None, workspace)
        if out is None:
            out = torch.empty(shape, dtype=out_dtype, device=dev)
        assert out.is_contiguous() and tuple(out.shape) == shape and out.dtype in (torch.int32, torch.int16)
        assemble_split(mask, vectors, scale, sparse, out)
        if not check:
            return out
        if not (int(sparse.status.item()) & L.STATUS_ROOT_OVERFLOW):
            if out.dtype == torch.int16 and sparse.num_components + 2 > 32767:  # same guard as the fused path below
                raise RuntimeError(f"{sparse.num_components} components do not fit int16 instance labels; use out_dtype=torch.int32")
            return out
        workspace = None  # more tile-local components than the default capacity: redo at the worst case, fused
    sparse = label_components(mask, planar=False, label_base=2, workspace=workspace, check=check)
    want16 = (out.dtype if out is not None else out_dtype) == torch.int16
    if check and want16 and sparse.num_components + 2 > 32767:
        # the reference's int16 instance mask wraps silently here (SURVEY B#5); say so instead
        raise RuntimeError(f"{sparse.num_components} components do not fit int16 instance labels; use out_dtype=torch.int32")
    return gather_instances(vectors, scale, sparse, N=N, decay=decay, crop=crop, overlap=overlap, out=out,
                            out_dtype=out_dtype)


class GraphedAssembler:
    """`assemble_instances` for SMALL volumes as one CUDA-graph launch.

    Below a few million voxels a pass is launch-bound: ~11 dependent kernels of a few microseconds each cost ~0.1 ms
    when issued one by one from Python (profiles/r01_rows.json: C1 128x128x32 in 0.099 ms = 0.9 % of HBM, a 300x300x20
    tile's flood fill in 0.155 ms).  The chain has no host decision inside (sizes are fixed by the shape, overflow is a
    status bit), so it is captured once — on static input / output buffers owned by this object — and replayed with a
    single launch.  `__call__` copies the inputs into the static buffers (device to device, or from the host) and
    replays; `check=True` reads the status word afterwards."""

    def __init__(self, shape: Tuple[int, int, int], scale, device="cuda:0", vec_dtype=torch.float16, N: int = 1, decay: float = 1.0,
                 crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0), out_dtype: torch.dtype = torch.int32):
        X, Y, Z = shape
        self.shape, self.dev = (X, Y, Z), torch.device(device)
        self.mask = torch.zeros((X, Y, Z), dtype=torch.uint8, device=self.dev)
        self.vec = torch.zeros((3, X, Y, Z), dtype=vec_dtype, device=self.dev)
        self.out = torch.empty((X, Y, Z), dtype=out_dtype, device=self.dev)
        self._args = dict(N=N, decay=decay, crop=crop, overlap=tuple(overlap))
        self._scale = as_floats(scale, 3)
        self.sparse = new_sparse(self.shape, self.dev)

        def chain():
            launch_label(self.mask, self.sparse, False, 2)
            gather_instances(self.vec, self._scale, self.sparse, out=self.out, **self._args)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            chain()  # warm-up outside the capture (first-launch work, the workspace's clean marker)
            torch.cuda.synchronize(self.dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                chain()
        torch.cuda.current_stream(self.dev).wait_stream(side)

    def __call__(self, skeleton_mask: Tensor, vectors: Tensor, check: bool = True) -> Tensor:
        m = skeleton_mask.squeeze(0) if skeleton_mask.ndim == 4 else skeleton_mask
        self.mask.copy_(m if m.dtype == torch.uint8 else m.gt(0), non_blocking=True)
        self.vec.copy_(vectors, non_blocking=True)
        self.graph.replay()
        if check:
            self.sparse.check()
            if self.out.dtype == torch.int16 and self.sparse.num_components + 2 > 32767:
                raise RuntimeError(f"{self.sparse.num_components} components do not fit int16 instance labels")
        return self.out


class HostAssembler:
    """End-to-end form of `assemble_instances` for volumes that live in HOST memory (the reference
    keeps them in zarr / numpy, skoots/lib/eval.py:102-103,223,245): copies the u8 skeleton mask
    and the fp16 vectors host->device, labels + gathers on the GPU, and copies the instance mask
    back.  Device buffers and the CCL workspace are allocated once and reused across calls.

    The pass is pipelined over X-slabs on three streams: the mask goes up first and is labelled while
    the vector field follows slab by slab; each slab is gathered as soon as its vectors have landed
    and its labels start travelling back while the next slab is still arriving, so the two PCIe
    directions and the kernels overlap (N = 1; with N > 1 a walk may read vectors of any slab of its
    crop, so the gather waits for the whole field).
    """

    def __init__(self, shape: Tuple[int, int, int], device="cuda:0", vec_dtype=torch.float16,
                 out_dtype=torch.int32, n_slabs: int = 16):
        X, Y, Z = shape
        self.shape, self.dev = (X, Y, Z), torch.device(device)
        self.mask = torch.empty((X, Y, Z), dtype=torch.uint8, device=self.dev)
        self.vec = torch.empty((3, X, Y, Z), dtype=vec_dtype, device=self.dev)
        self.out = torch.empty((X, Y, Z), dtype=out_dtype, device=self.dev)
        self.workspace = None
        self.up = torch.cuda.Stream(self.dev)
        self.down = torch.cuda.Stream(self.dev)
        # slab starts must fall on multiples of 256 voxels (skb_assemble_range)
        plane = Y * Z
        bounds = sorted({X * i // n_slabs for i in range(n_slabs + 1)})
        self.bounds = [b for b in bounds if (b * plane) % 256 == 0 or b == X]
        if self.bounds[0] != 0:
            self.bounds = [0] + self.bounds

    def __call__(self, mask_host: Tensor, vec_host: Tensor, scale, out_host: Tensor, N: int = 1, decay: float = 1.0,
                 crop=None, overlap=(0, 0, 0), check: bool = True) -> Tensor:
        X, Y, Z = self.shape
        plane = Y * Z
        main = torch.cuda.current_stream(self.dev)
        self.up.wait_stream(main)
        self.down.wait_stream(main)
        mask_host, vec_host, out_host = mask_host.reshape(X, Y, Z), vec_host.reshape(3, X, Y, Z), out_host.reshape(X, Y, Z)
        slabs = list(zip(self.bounds[:-1], self.bounds[1:])) if N == 1 else [(0, X)]
        landed = []
        with torch.cuda.stream(self.up):
            self.mask.copy_(mask_host, non_blocking=True)
            mask_ready = torch.cuda.Event()
            mask_ready.record(self.up)
            for x0, x1 in slabs:
                for c in range(3):  # one contiguous block per channel: a strided slice would not go out as a plain DMA
                    self.vec[c, x0:x1].copy_(vec_host[c, x0:x1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.up)
                landed.append(ev)
        main.wait_event(mask_ready)
        sparse = label_components(self.mask, label_base=2, workspace=self.workspace, check=False)
        self.workspace = sparse.workspace
        for (x0, x1), ev in zip(slabs, landed):
            main.wait_event(ev)
            gather_instances(self.vec, scale, sparse, N=N, decay=decay, crop=crop, overlap=overlap, out=self.out,
                             voxel_range=(x0 * plane, (x1 - x0) * plane))
            done = torch.cuda.Event()
            done.record(main)
            self.down.wait_event(done)
            with torch.cuda.stream(self.down):
                out_host[x0:x1].copy_(self.out[x0:x1], non_blocking=True)
        torch.cuda.synchronize(self.dev)
        if check:
            sparse.check()
            if self.out.dtype == torch.int16 and sparse.num_components + 2 > 32767:
                raise RuntimeError(f"{sparse.num_components} components do not fit int16 instance labels; use out_dtype=torch.int32")
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


@torch.inference_mode()
def segment_volume(model, image: Tensor, vector_scale, dataset_mean: float, dataset_std: float,
                   tile: Sequence[int] = (300, 300, 20), tile_overlap: Sequence[int] = (50, 50, 5),
                   threshold: float = 0.8, N: int = EVAL_N, decay: float = 1.0,
                   crop: Sequence[int] = EVAL_CROP, overlap: Sequence[int] = EVAL_OVERLAP,
                   device="cuda:0", autocast: bool = True) -> Tensor:
    """The compute of `skoots.lib.eval.eval` (skoots/lib/eval.py:126-176 and 223-284) with every
    intermediate kept on the device: tile the image, run `model` on each tile, fuse the tile epilogue
    straight into device-resident whole-volume vectors / skeleton arrays (no `.cpu().numpy()` -> zarr
    round trip), then label + assemble.  `model` is the caller's network (the reference's
    bism UNet): a callable mapping (1,1,x,y,z) float -> (1,C>=5,x,y,z).  image: (C,X,Y,Z), CPU or CUDA.
    Returns the int16 instance mask (X,Y,Z) on the device, before `fastremap.renumber`."""
    from .lib.cropper import crops
    dev = torch.device(device)
    c, X, Y, Z = image.shape
    vectors = torch.zeros((3, X, Y, Z), dtype=torch.float16, device=dev)
    skeleton = torch.zeros((1, X, Y, Z), dtype=torch.uint8, device=dev)
    tile = list(tile)
    for piece, (x, y, z) in crops(image, tile, tuple(tile_overlap), device=dev):
        piece = piece.sub(dataset_mean).div(dataset_std)
        with torch.autocast("cuda", enabled=autocast):
            out = model(piece.float())
        tile_epilogue(out, vectors, skeleton, (x, y, z), tile_overlap, threshold)
    return assemble_instances(skeleton, vectors, vector_scale, N=N, decay=decay, crop=crop, overlap=overlap,
                              out_dtype=torch.int16)
