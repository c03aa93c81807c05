"""What the reference does to an instance mask after assembly, on B200:
`fastremap.renumber` (skoots/lib/eval.py:304) and the per-object validation metrics of
`skoots/validate/lib.py` — same names, arguments and return conventions.  CUDA tensors are
processed where they live; host tensors (what `skoots-validate` reads from tif files) are staged through the GPU.

The reference's `mask_iou` / `mask_dice` loop over every (gt object, predicted object) pair with full-volume
boolean passes — O(N·M·V).  Here one pass over the two masks fills a contingency table (`skb_contingency`)
and a second kernel turns counts into ratios; the numbers are the same fp32 values.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L


def _as_labels(t: Tensor) -> Tensor:
    L.require_cuda(t)
    if t.dtype not in (torch.int16, torch.int32):
        if t.dtype in (torch.int64, torch.uint8, torch.int8, torch.bool):
            t = t.to(torch.int32)
        else:
            raise L.SkootsB200Error(f"instance masks must hold integers, got {t.dtype}")
    return t.contiguous()


def label_max(labels: Tensor) -> int:
    """largest label of an integer CUDA volume (one small kernel + a 4-byte read)."""
    labels = _as_labels(labels)
    out = torch.empty(1, dtype=torch.int32, device=labels.device)
    with torch.cuda.device(labels.device):
        L.check(L.load().skb_label_max(labels.data_ptr(), L.dtype_code(labels), labels.numel(), out.data_ptr(),
                                       L.stream_ptr(labels.device)))
    return int(out.item())


def _check_status(status: Tensor, who: str) -> None:
    if int(status.item()) & L.STATUS_LABEL_RANGE:
        raise L.SkootsB200Error(f"{who}: a label is negative or not smaller than the table size")


def renumber(arr: Tensor, in_place: bool = False, max_label: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """`fastremap.renumber(arr, in_place=...)` (skoots/lib/eval.py:304) for int16 / int32 CUDA masks: labels become
    1..N in order of first appearance in the C-order scan, 0 stays 0.  Returns (renumbered, remap) where
    remap[old] = new (0 for labels that do not occur) — fastremap returns the same mapping as a dict and may also
    shrink the dtype, which this does not."""
    dev, staged = L.compute_device(arr)
    if staged:  # eval.py:304 renumbers a host array: up, renumber, back (into the same storage when in_place)
        out, remap = renumber(L.stage_in(arr, dev), in_place=False, max_label=max_label)
        if in_place:
            arr.copy_(out)
            return arr, remap.cpu()
        return out.cpu(), remap.cpu()
    src = _as_labels(arr)
    if src.numel() == 0:
        return (arr if in_place else arr.clone()), torch.zeros(1, dtype=torch.int32, device=arr.device)
    if in_place and (src.data_ptr() != arr.data_ptr()):
        raise L.SkootsB200Error("renumber(in_place=True) needs a contiguous int16 / int32 CUDA tensor")
    out = src if in_place else src.clone()
    dev = out.device
    table = (label_max(out) if max_label is None else int(max_label)) + 1
    table = max(table, 2)
    lib = L.load()
    ws = torch.empty(lib.skb_renumber_workspace_bytes(out.numel(), table), dtype=torch.uint8, device=dev)
    remap = torch.empty(table, dtype=torch.int32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)  # [n_labels, status]
    with torch.cuda.device(dev):
        L.check(lib.skb_renumber(out.data_ptr(), L.dtype_code(out), out.numel(), table, ws.data_ptr(), ws.numel(),
                                 remap.data_ptr(), meta[0:1].data_ptr(), meta[1:2].data_ptr(), L.stream_ptr(dev)))
    _check_status(meta[1], "renumber")
    return out.view(arr.shape), remap


class _Contingency:
    """intersection / area counts of every (gt object, predicted object) pair, objects in sorted label order."""

    def __init__(self, gt: Tensor, pred: Tensor):
        assert gt.shape == pred.shape, "Input tensors must be the same shape"        # validate/lib.py:198
        assert gt.device == pred.device, "Input tensors must be on the same device"  # validate/lib.py:199
        dev, self.staged = L.compute_device(gt, pred)  # skoots-validate passes host tensors read from tif files
        if self.staged:
            gt, pred = L.stage_in(gt, dev), L.stage_in(pred, dev)
        gt, pred = _as_labels(gt), _as_labels(pred)
        dev, lib, n = gt.device, L.load(), gt.numel()
        self.dev = dev
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        idx, counts, self.values = [], [], []
        with torch.cuda.device(dev):
            for vol in (gt, pred):
                table = max(label_max(vol) + 1, 2)
                index = torch.empty(table, dtype=torch.int32, device=dev)
                values = torch.empty(table, dtype=torch.int32, device=dev)
                count = torch.zeros(1, dtype=torch.int32, device=dev)
                ws = torch.empty(lib.skb_unique_index_workspace_bytes(table), dtype=torch.uint8, device=dev)
                L.check(lib.skb_unique_index(vol.data_ptr(), L.dtype_code(vol), n, table, index.data_ptr(), values.data_ptr(),
                                             count.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(), L.stream_ptr(dev)))
                idx.append(index)
                counts.append(count)
                self.values.append(values)
            self.N, self.M = int(counts[0].item()), int(counts[1].item())
            self.values = [self.values[0][:self.N], self.values[1][:self.M]]
            self.inter = torch.zeros((self.N, self.M), dtype=torch.int32, device=dev)
            self.area_gt = torch.zeros(max(self.N, 1), dtype=torch.int32, device=dev)
            self.area_pred = torch.zeros(max(self.M, 1), dtype=torch.int32, device=dev)
            if self.N and self.M:
                L.check(lib.skb_contingency(gt.data_ptr(), L.dtype_code(gt), pred.data_ptr(), L.dtype_code(pred), n,
                                            idx[0].data_ptr(), idx[0].numel(), idx[1].data_ptr(), idx[1].numel(), self.N, self.M,
                                            self.inter.data_ptr(), self.area_gt.data_ptr(), self.area_pred.data_ptr(),
                                            L.stream_ptr(dev)))
        _check_status(status[0], "mask_iou / mask_dice")

    def ratios(self, want_iou: bool, want_dice: bool):
        iou = torch.zeros((self.N, self.M), dtype=torch.float32, device=self.dev) if want_iou else None
        dice = torch.zeros((self.N, self.M), dtype=torch.float32, device=self.dev) if want_dice else None
        if self.N and self.M:
            with torch.cuda.device(self.dev):
                L.check(L.load().skb_iou_dice(self.inter.data_ptr(), self.area_gt.data_ptr(), self.area_pred.data_ptr(), self.N,
                                              self.M, L.ptr(iou), L.ptr(dice), L.stream_ptr(self.dev)))
        return iou, dice


def mask_iou(gt: Tensor, pred: Tensor) -> Tensor:
    """skoots/validate/lib.py:190-229 — N x M matrix of IoUs, rows = sorted gt labels > 0, columns = sorted
    predicted labels > 0, float32; 0 where two objects do not touch."""
    c = _Contingency(gt, pred)
    return L.stage_out(c.ratios(True, False)[0], c.staged)


def mask_dice(gt: Tensor, pred: Tensor) -> Tensor:
    """skoots/validate/lib.py:232-275 — N x M matrix of Dice indices.  Like the reference (its assert at :266-268)
    this raises AssertionError when a ground-truth object and a predicted object coincide exactly (dice == 1)."""
    c = _Contingency(gt, pred)
    dice = c.ratios(False, True)[1]
    if dice.numel():
        both = c.area_gt[:c.N, None].to(torch.int64) + c.area_pred[None, :c.M].to(torch.int64)
        assert not bool((2 * c.inter.to(torch.int64) >= both).logical_and(c.inter > 0).any()), "numerator >= denominator"
    return L.stage_out(dice, c.staged)


def accuracies_from_iou(iou: Tensor, thr: float = 0.1) -> Tuple[int, int, int]:
    """skoots/validate/lib.py:170-187 — (true positives, false positives, false negatives) at an IoU threshold."""
    dev, staged = L.compute_device(iou)
    if staged:
        iou = L.stage_in(iou, dev)
    if iou.ndim != 2 or iou.shape[0] == 0 or iou.shape[1] == 0:
        # the reference's iou.max(dim=1) raises on an empty dimension
        raise IndexError("accuracies_from_iou: the IoU matrix must have at least one row and one column")
    iou = iou.contiguous().float()
    n, m = iou.shape
    scratch = torch.empty(n + m, dtype=torch.int32, device=iou.device)
    out = torch.empty(3, dtype=torch.int32, device=iou.device)
    with torch.cuda.device(iou.device):
        L.check(L.load().skb_accuracies_from_iou(iou.data_ptr(), n, m, float(thr), scratch.data_ptr(), out.data_ptr(),
                                                 L.stream_ptr(iou.device)))
    tp, fp, fn = (int(v) for v in out.tolist())
    return tp, fp, fn
