"""TEST INFRASTRUCTURE — checks a piece of a FULL-VOLUME GPU result against the CPU reference.

bench.py's 2048x2048x512 output cannot be recomputed by the CPU path in a bench run (minutes at N = 1, an hour at
N = 10), and a sub-volume re-run is a different problem (other clamps, other component numbering).  So:

  R = the first Rx x Ry x Rz voxels of the workload (a box the CPU path labels in seconds; for eval mode its crop grid
      coincides with the full volume's grid over the compared part);
  S = R minus a margin on every face R cuts through the volume — wide enough that no voxel of S can reach, with its
      vector(s), a place where the sub-volume clamps differently from the full volume;

the CPU path (the unmodified reference via oracle/ref_runner.py when it is importable, else the oracle port) runs on
R; the GPU's FULL-VOLUME output restricted to S must then equal the CPU result on S up to the canonical relabelling
north_star allows (component numbers are raster ranks, which differ between R and the whole volume) — checked as
"same zero pattern and a one-to-one correspondence between the two label sets" — except for voxels whose component
touches a face where R cuts the volume (the sub-volume run sees such a component truncated, possibly in pieces; every
piece touches that face, so they are identified exactly and excluded; the count is reported).
"""
import os
import sys
import time
from typing import Dict, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

EVAL_CROP, EVAL_OVERLAP = (500, 500, 50), (50, 50, 5)  # skoots/lib/eval.py:248-249

# boxes tried in order: the first whose estimated CPU time fits the budget is used (shape of R, per-axis margin)
_LADDER_WHOLE = [((960, 960, 160), (128, 128, 32)), ((704, 704, 160), (128, 128, 32)), ((512, 512, 160), (128, 128, 32)),
                 ((384, 384, 96), (96, 96, 32)), ((256, 256, 64), (64, 64, 16)), ((128, 128, 32), (0, 0, 0))]
_LADDER_EVAL = [((900, 900, 90), (100, 100, 30)), ((500, 900, 90), (100, 100, 30)), ((500, 500, 90), (100, 100, 30)),
                ((500, 500, 50), (100, 100, 20)), ((128, 128, 32), (0, 0, 0))]
# rough CPU throughput of the reference path in voxels/s per hop count, only to size the box
_EST_RATE = {"whole": 20e6, "eval": 18e6}


def _owners(dim: int, size: int, ov: int):
    """owner crop origin of every coordinate of an axis under eval()'s loop (later crops overwrite), -1 = never written"""
    size = min(size, dim)
    own = np.full(dim, -1, dtype=np.int64)
    o = 0
    while o < dim:  # cropper.py:100-142: advance by size - 2*ov, last crop shifted back to dim - size
        origin = o if o + size <= dim else dim - size
        own[origin + ov:origin + size - ov] = origin
        o += size - 2 * ov
    return own


def _common_prefix(dim_r: int, dim_full: int, size: int, ov: int) -> int:
    """how many leading coordinates are owned by the same crop in a dim_r-long sub-volume and in the full axis"""
    a, b = _owners(dim_r, size, ov), _owners(dim_full, size, ov)[:dim_r]
    diff = np.nonzero(a != b)[0]
    return int(diff[0]) if len(diff) else dim_r


def plan(shape, mode: str, hops: int, budget_s: float, ladder=None) -> Dict:
    """the sample box for a workload: {"R": (x,y,z), "S": (x,y,z), "text": ...}; R and S start at the volume's origin."""
    ladder = ladder or (_LADDER_EVAL if mode == "eval" else _LADDER_WHOLE)
    rate = _EST_RATE[mode] / (1.0 if hops == 1 else 0.45 * hops)
    pick = ladder[-1]
    for R, margin in ladder:
        Rc = tuple(min(r, d) for r, d in zip(R, shape))
        if Rc[0] * Rc[1] * Rc[2] / rate <= budget_s:
            pick = (R, margin)
            break
    R, margin = pick
    R = tuple(min(r, d) for r, d in zip(R, shape))
    S = tuple(r if r == d else max(1, r - m) for r, d, m in zip(R, shape, margin))
    if mode == "eval":  # compare only where the sub-volume's crop grid IS the full volume's grid
        S = tuple(min(s_, _common_prefix(r, d, c, o)) if r != d else s_
                  for s_, r, d, c, o in zip(S, R, shape, EVAL_CROP, EVAL_OVERLAP))
    return {"R": R, "S": S,
            "text": f"first {R[0]}x{R[1]}x{R[2]} voxels of the workload volume, one whole pass per step"}


def run_cpu(mask_R: torch.Tensor, vec_R: torch.Tensor, scale, hops: int, mode: str) -> Tuple[torch.Tensor, torch.Tensor, float, str]:
    """(instance labels on R, skeleton labels on R, seconds, kind) — kind "reference" = the unmodified reference's
    functions (oracle/ref_runner.py), "port" = oracle/skoots_oracle.py when the reference is not importable."""
    crop, overlap = (EVAL_CROP, EVAL_OVERLAP) if mode == "eval" else (None, (0, 0, 0))
    out_dtype = torch.int16 if mode == "eval" else torch.int32
    scale = scale if isinstance(scale, torch.Tensor) else torch.tensor(scale)
    import ref_runner
    if ref_runner.available():
        ref_runner.load()
        t0 = time.perf_counter()
        labels = ref_runner.flood_fill(mask_R)
        inst = ref_runner.assemble(labels, vec_R, scale, hops, 1.0, crop, overlap, out_dtype)
        return inst, labels, time.perf_counter() - t0, "reference"
    import skoots_oracle as orc
    t0 = time.perf_counter()
    labels = orc.flood_fill_exact(mask_R.to(torch.int16))
    dims = tuple(labels.shape)
    inst = orc.assemble_instances(labels, vec_R, scale, N=hops, crop=dims if crop is None else crop, overlap=overlap,
                                  out_dtype=out_dtype)
    return inst, labels, time.perf_counter() - t0, "port"


def compare(got_S, want_R, labels_R, S, full_shape) -> Dict:
    """got_S: the GPU's full-volume output restricted to the S box (array-like, any int dtype)."""
    g = np.ascontiguousarray(got_S.cpu().numpy() if isinstance(got_S, torch.Tensor) else got_S).astype(np.int64)
    R = tuple(want_R.shape)
    w = want_R[:S[0], :S[1], :S[2]].numpy().astype(np.int64)
    lab = labels_R.numpy()
    cut = set()
    for axis in range(3):
        if R[axis] < full_shape[axis]:  # this face of R runs through the volume
            cut.update(np.unique(np.take(lab, R[axis] - 1, axis=axis)).tolist())
    cut.discard(0)
    keep = np.ones(w.shape, dtype=bool)
    if cut:
        keep &= ~np.isin(w, np.fromiter(cut, dtype=np.int64))
    excluded = int(keep.size - keep.sum())
    gz, wz = (g == 0), (w == 0)
    zero_mismatch = int(((gz != wz) & keep).sum())
    nz = keep & ~gz & ~wz
    key = np.unique(g[nz] * (1 << 32) + w[nz])
    gs, ws = key >> 32, key & 0xFFFFFFFF
    one_to_one = len(np.unique(gs)) == len(gs) and len(np.unique(ws)) == len(ws)
    ok = zero_mismatch == 0 and one_to_one
    return {"ok": bool(ok), "compared_voxels": int(keep.sum()), "labelled_compared": int(nz.sum()), "labels_compared": int(len(key)),
            "excluded_voxels_of_components_cut_by_the_sample_box": excluded, "zero_pattern_mismatches": zero_mismatch,
            "one_to_one": bool(one_to_one)}
