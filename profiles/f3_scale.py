#!/usr/bin/env python
"""Row f3 at scale: `efficient_flood_fill(..., reference_crops=True)` — the reference's crop-by-crop labelling with its
seam heuristic (skoots/lib/flood_fill.py:27-122) — on the bench workload volume (2048 x 2048 x 512: 3 x 3 x 3 crops of
1000 x 1000 x 200), next to the default exact labelling of the same volume, and checked against the oracle's CPU
restatement of the reference on a volume the CPU finishes in seconds.

    python profiles/f3_scale.py > profiles/r02_f3_scale.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

from skoots_b200.lib.flood_fill import efficient_flood_fill
from skoots_b200.synthetic import make_tube_volume

DEV = "cuda:0"


def timed(fn, src, iters=3):
    best, out = 1e30, None
    for _ in range(iters):
        vol = src.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(vol)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


def main():
    rows = []
    # parity at a size the CPU restatement handles: 2 x 2 x 2 crops
    import skoots_oracle as orc
    shape = (1100, 1100, 256)
    tv = make_tube_volume(shape, 1200, seed=3, device=DEV, want_mask=False, want_skeleton_dict=False)
    src = tv.skeleton.to(torch.int16)
    ms, got = timed(lambda v: efficient_flood_fill(v, reference_crops=True), src, iters=2)
    t0 = time.perf_counter()
    want = orc.flood_fill_multicrop(src.cpu().clone())
    cpu_ms = (time.perf_counter() - t0) * 1e3
    rows.append({"row": "f3 reference_crops=True vs the oracle's restatement of flood_fill.py:27-122", "shape": list(shape),
                 "bit_identical": bool(np.array_equal(got.cpu().numpy(), want.numpy())), "labels": int(got.max()),
                 "gpu_ms": round(ms, 2), "cpu_oracle_ms": round(cpu_ms, 1)})
    del tv, src, got, want
    torch.cuda.empty_cache()

    # the bench workload's shape.  With its 16 384 tubes the crop-local labels (every tube piece in every crop, plus one
    # skipped number per crop) exceed 32 767: the reference's int16 volume wraps there, this implementation raises
    # RuntimeError (measured: profiles/r02_f3_scale.json of the first run).  8 000 tubes fit.
    shape = (2048, 2048, 512)
    tv = make_tube_volume(shape, 8000, seed=0, device=DEV, want_mask=False, want_skeleton_dict=False)
    src = tv.skeleton.to(torch.int16)
    del tv
    exact_ms, exact = timed(lambda v: efficient_flood_fill(v), src)
    try:
        quirk_ms, quirk = timed(lambda v: efficient_flood_fill(v, reference_crops=True), src)
        same_fg = bool(torch.equal(exact > 0, quirk > 0))
        # the reference's heuristic can only merge more than the exact labelling does: every exact component maps to ONE label
        pairs = torch.unique(torch.stack([exact[exact > 0].int(), quirk[quirk > 0].int()]), dim=1)
        rows.append({"row": "f3 reference_crops=True at the bench workload's shape, 8000 tubes (27 crops, 18 seam planes)", "shape": list(shape),
                     "gpu_ms": round(quirk_ms, 1), "exact_labelling_ms": round(exact_ms, 2), "same_foreground": same_fg,
                     "exact_components": int(torch.unique(pairs[0]).numel()), "reference_crops_labels": int(torch.unique(pairs[1]).numel()),
                     "exact_component_split_over_labels": int(pairs.shape[1] - torch.unique(pairs[0]).numel()),
                     "note": "per-crop labelling, seam test (presence tables) and the final table lookup on the GPU; host part = the walk over "
                             "the label graph and one small read per seam"})
    except RuntimeError as exc:  # more than 32767 crop-local labels: the reference's int16 volume overflows there too
        rows.append({"row": "f3 reference_crops=True at the bench workload size", "shape": list(shape), "exact_labelling_ms": round(exact_ms, 2),
                     "error": str(exc)[:200]})
    print(json.dumps({"rows": rows}, indent=1))


if __name__ == "__main__":
    main()
