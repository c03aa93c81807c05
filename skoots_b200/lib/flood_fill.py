"""`skoots.lib.flood_fill.efficient_flood_fill` on B200 (reference: skoots/lib/flood_fill.py)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L


class SparseLabels:
    """Result of the sparse CCL: the workspace the fused gather reads, plus counters."""

    def __init__(self, workspace: Tensor, shape: Tuple[int, int, int], ncomp: Tensor, status: Tensor, capacity: int):
        self.workspace, self.shape, self.ncomp, self.status, self.capacity = workspace, shape, ncomp, status, capacity

    def check(self) -> None:
        """Raises if the device-side status word reports an overflow (one 4-byte D2H read)."""
        if int(self.status.item()) & L.STATUS_ROOT_OVERFLOW:
            raise L.SkootsB200Error("CCL workspace overflow: more tile-local components than capacity")

    @property
    def num_components(self) -> int:
        return int(self.ncomp.item())


def default_capacity(voxels: int) -> int:
    # tile-local components are bounded by V/2+1 (checkerboard); skeleton masks are ~1 % foreground,
    # so V/8 is already generous.  label_components() retries at the worst case on overflow.
    return max(1 << 16, voxels // 8)


def _as_mask(mask: Tensor) -> Tensor:
    if mask.ndim != 3:
        raise RuntimeError(f"mask must be (X,Y,Z), got {tuple(mask.shape)}")
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    elif mask.dtype not in (torch.uint8, torch.int16):
        mask = mask.gt(0).view(torch.uint8)
    return mask.contiguous()


def _workspace_for(shape, cap: int, workspace: Optional[Tensor], dev) -> Tensor:
    need = L.load().skb_ccl_workspace_bytes(*shape, cap)
    if workspace is None or workspace.numel() < need or workspace.device != dev:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    return workspace


def launch_label(mask: Tensor, sparse: "SparseLabels", planar: bool, label_base: int, phase: int = 0, tiles: int = 0) -> None:
    """enqueues the labelling of `mask` into sparse.workspace on the current stream.  phase = 0 (all),
    L.CCL_PHASE_PACK (header + bit mask only) or L.CCL_PHASE_LABEL (everything after the pack);
    tiles = 0 | L.CCL_TILES_DYNAMIC | L.CCL_TILES_STATIC forces the tile kernel's work distribution."""
    dev = mask.device
    X, Y, Z = sparse.shape
    ws = sparse.workspace
    flags = phase | tiles
    if phase != L.CCL_PHASE_LABEL:
        # a workspace that already went through a pass of this shape has its root bitmap zeroed (see the C header)
        key = (X, Y, Z, sparse.capacity, ws.data_ptr())
        if getattr(ws, "_skb_clean", None) == key:
            flags |= L.CCL_WORKSPACE_CLEAN
        ws._skb_clean = None
    with torch.cuda.device(dev):
        L.check(L.load().skb_ccl_label_sparse(mask.data_ptr(), L.dtype_code(mask), X, Y, Z, int(planar), int(label_base),
                                              sparse.capacity, ws.data_ptr(), ws.numel(), sparse.ncomp.data_ptr(),
                                              sparse.status.data_ptr(), flags, L.stream_ptr(dev)))
    if phase != L.CCL_PHASE_PACK:
        ws._skb_clean = (X, Y, Z, sparse.capacity, ws.data_ptr())


def new_sparse(shape, dev, capacity: Optional[int] = None, workspace: Optional[Tensor] = None) -> "SparseLabels":
    X, Y, Z = shape
    cap = int(capacity) if capacity else default_capacity(X * Y * Z)
    workspace = _workspace_for((X, Y, Z), cap, workspace, dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    return SparseLabels(workspace, (X, Y, Z), meta[0], meta[1], cap)


def label_components(mask: Tensor, planar: bool = False, label_base: int = 2, capacity: Optional[int] = None,
                     workspace: Optional[Tensor] = None, check: bool = True) -> SparseLabels:
    """Connected components of `mask > 0` in the sparse on-device form.
    mask: (X,Y,Z) uint8/bool/int16 CUDA tensor.  6-connectivity, or per-x-plane 4-connectivity when
    `planar`.  Labels are label_base+1.. in scipy.ndimage.label's raster order."""
    dev = L.require_cuda(mask)
    mask = _as_mask(mask)
    X, Y, Z = mask.shape
    V = X * Y * Z
    cap = int(capacity) if capacity else default_capacity(V)
    while True:
        res = new_sparse((X, Y, Z), dev, cap, workspace)
        launch_label(mask, res, planar, label_base)
        if not check:
            return res
        if int(res.status.item()) & L.STATUS_ROOT_OVERFLOW and cap < V // 2 + 1:
            cap, workspace = V // 2 + 1, None
            continue
        res.check()
        return res


def write_dense(sparse: SparseLabels, out: Tensor) -> Tensor:
    dev = L.require_cuda(out)
    X, Y, Z = sparse.shape
    assert out.is_contiguous() and out.numel() == X * Y * Z and out.dtype in (torch.int16, torch.int32)
    with torch.cuda.device(dev):
        L.check(L.load().skb_ccl_write_dense(sparse.workspace.data_ptr(), X, Y, Z, out.data_ptr(), L.dtype_code(out),
                                             L.stream_ptr(dev)))
    return out


REFERENCE_CROP = (1000, 1000, 200)  # skoots/lib/flood_fill.py:28


def _wrap16(v: Tensor) -> Tensor:
    """int32 -> the value int16 arithmetic would have produced (two's complement wrap), still int32."""
    return torch.remainder(v + 32768, 65536) - 32768


def _adjacent_by_sum_product(p0: Tensor, p1: Tensor):
    """get_adjacent_labels, flood_fill.py:237-261, on two seam planes (device int16 tensors): labels (a, b) count as
    touching when a+b and a*b (int16 arithmetic) both occur among the element-wise sums / products of the planes.
    On the device: the sums, products and labels that occur are marked in 65 536-entry presence tables (one scatter
    each), the reference's double Python loop over the label pairs is one outer sum / product looked up in them.  One
    small D2H read (the pairs), same pairs in the same order: a ascending, then b ascending."""
    a32, b32 = p0.reshape(-1).to(torch.int32), p1.reshape(-1).to(torch.int32)

    def present(values):
        table = torch.zeros(65536, dtype=torch.bool, device=values.device)
        table[(values + 32768).long()] = True
        return table

    sums, prods = present(_wrap16(a32 + b32)), present(_wrap16(a32 * b32))
    a = (present(a32).nonzero().squeeze(1) - 32768).to(torch.int32)
    b = (present(b32).nonzero().squeeze(1) - 32768).to(torch.int32)
    a, b = a[a != 0], b[b != 0]
    if a.numel() == 0 or b.numel() == 0:
        return []
    hit = sums[(_wrap16(a[:, None] + b[None, :]) + 32768).long()] & prods[(_wrap16(a[:, None] * b[None, :]) + 32768).long()]
    ia, ib = hit.nonzero(as_tuple=True)
    return list(zip(a[ia].tolist(), b[ib].tolist()))


def _flood_fill_reference_crops(vol: Tensor, crop=REFERENCE_CROP) -> Tensor:
    """Row f3: the reference's efficient_flood_fill INCLUDING its multi-crop behaviour (flood_fill.py:27-122), for
    volumes larger than one 1000x1000x200 crop: every crop is labelled on its own (numbering continues from the
    previous crop's maximum — and restarts after an empty crop, :140), labels that meet at a crop seam are found
    with the sum/product test and every group of them is replaced by its last-visited member.  The per-crop
    labelling, the seam test and the replacement run on the GPU; only the walk over the (small) label graph is host
    logic.  Bit-identical to the reference on its own fixture (tests/golden/flood_multicrop.npz)."""
    import numpy as np
    from .cropper import _origins
    X, Y, Z = vol.shape
    size = [min(c, d) for c, d in zip(crop, (X, Y, Z))]
    max_id = 1
    seams = ([], [], [])
    for x in _origins(X, size[0], 0):
        for y in _origins(Y, size[1], 0):
            for z in _origins(Z, size[2], 0):
                for ax, o in enumerate((x, y, z)):
                    if o not in seams[ax]:
                        seams[ax].append(o)
                view = vol[x:x + size[0], y:y + size[1], z:z + size[2]]
                piece = view.contiguous()
                # flood_all(crop, max_id + 1), flood_fill.py:125-140: scipy's label (1..n) + (max_id + 1) on foreground
                sparse = label_components(piece, planar=False, label_base=max_id + 1)
                n = sparse.num_components
                if max_id + 1 + n > 32767:
                    raise RuntimeError("more labels than the reference's int16 volume can hold")
                write_dense(sparse, piece)
                view.copy_(piece)
                max_id = max_id + 1 + n if n else 0  # `mask.max()`: 0 for an empty crop, so numbering restarts (:140)
    pairs = []
    for ax in range(3):
        for o in seams[ax]:
            if o > 0:
                pairs.extend(_adjacent_by_sum_product(vol.select(ax, o), vol.select(ax, o - 1)))
    graph = {}
    for a, b in pairs:
        graph.setdefault(a, []).append(b)
        graph.setdefault(b, []).append(a)
    seen, table = set(), {}
    for start in graph:  # connected_components + dfs, flood_fill.py:143-174, iteratively
        if start in seen:
            continue
        order, stack = [start], [(start, iter(graph[start]))]
        seen.add(start)
        while stack:
            node, it = stack[-1]
            nxt = next((m for m in it if m not in seen), None)
            if nxt is None:
                stack.pop()
            else:
                seen.add(nxt)
                order.append(nxt)
                stack.append((nxt, iter(graph[nxt])))
        for member in order[:-1]:
            table.setdefault(member, order[-1])  # replaced by the LAST member (:100-104); first match wins (:197-203)
    if table:
        top = max(max(table), max(table.values())) + 1
        lut = np.arange(top, dtype=np.int32)
        for k, v in table.items():
            lut[k] = v
        lut_d = torch.from_numpy(lut).to(vol.device)
        with torch.cuda.device(vol.device):
            L.check(L.load().skb_apply_label_table(vol.data_ptr(), L.dtype_code(vol), vol.numel(), lut_d.data_ptr(), top,
                                                   L.stream_ptr(vol.device)))
    return vol


def efficient_flood_fill(skeleton: Tensor, reference_crops: bool = False) -> Tensor:
    """Drop-in for skoots.lib.flood_fill.efficient_flood_fill (:13-122): labels the connected
    components of `skeleton > 0` IN PLACE (int16, same storage) and returns the (X,Y,Z) view.

    Numbering: 3..N+2 in raster order of each component's first voxel — bit-identical to the
    reference for any volume that fits one of its 1000x1000x200 crops.  For larger volumes the
    reference labels crop by crop and merges at the seams with a heuristic (flood_fill.py:237-261) that
    can over-merge and re-use labels: by default this implementation returns the exact
    components instead; `reference_crops=True` reproduces the reference's result bit for bit, quirks included.
    """
    assert skeleton.dtype == torch.int16, f"Input tensor datatype must be int16 not {skeleton.dtype}"
    vol = skeleton.squeeze(0) if skeleton.ndim == 4 else skeleton
    if not vol.is_contiguous():
        raise RuntimeError("efficient_flood_fill labels in place and needs a contiguous tensor")
    dev, staged = L.compute_device(vol)
    if staged:  # eval() labels a HOST tensor (eval.py:223): up, label on the GPU, back into the same storage
        vol.copy_(efficient_flood_fill(L.stage_in(vol, dev), reference_crops))
        return vol
    if reference_crops and any(d > c for d, c in zip(vol.shape, REFERENCE_CROP)):
        return _flood_fill_reference_crops(vol)
    sparse = label_components(vol, planar=False, label_base=2)
    if sparse.num_components + 2 > 32767:
        raise RuntimeError(f"{sparse.num_components} components do not fit the reference's int16 labels")
    write_dense(sparse, vol)
    return vol
