"""TEST INFRASTRUCTURE — tests/golden/validate_metrics.npz: outputs of the UNMODIFIED reference functions
`skoots.validate.lib.{mask_iou, mask_dice, accuracies_from_iou}` (imported through oracle/ref_shim.py) on small
seeded instance masks.  Pins the oracle's restatement (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_validate.py) for row f2.

    PYTORCH_JIT=0 python oracle/gen_golden_validate.py
"""
import contextlib
import io
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch

import ref_shim

ref_shim.install()
import skoots.validate.lib as ref  # noqa: E402

from skoots_b200.synthetic import make_tube_volume  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    # ground truth = analytic tube ids; prediction = the same objects shifted, two of them merged, one split, one missing
    tv = make_tube_volume((48, 40, 16), 14, seed=21)
    gt = tv.mask.to(torch.int32)
    pred = torch.roll(gt, shifts=(1, -1, 0), dims=(0, 1, 2)).clone()
    ids = [int(v) for v in torch.unique(gt).tolist() if v != 0]
    pred[pred == ids[0]] = ids[1]                      # merge
    sel = pred == ids[2]
    half = torch.zeros_like(sel)
    half[: sel.shape[0] // 2] = True
    pred[sel & half] = 900                             # split
    pred[pred == ids[3]] = 0                           # missing
    pred[2:6, 30:34, 10:13] = 1200                     # a false positive far from everything
    iou = quiet(ref.mask_iou, gt, pred)
    dice = quiet(ref.mask_dice, gt, pred)
    pack = dict(gt=gt.numpy(), pred=pred.numpy(), iou=iou.numpy(), dice=dice.numpy())
    for thr in (0.1, 0.3, 0.5, 0.75):
        pack[f"acc_{int(thr * 100)}"] = np.array(ref.accuracies_from_iou(iou, thr), dtype=np.int64)
    # int16 masks with sparse label values
    g = torch.Generator().manual_seed(5)
    a = (torch.randint(0, 6, (12, 10, 8), generator=g) * 7).to(torch.int16)
    b = (torch.randint(0, 5, (12, 10, 8), generator=g) * 11).to(torch.int16)
    pack.update(gt_small=a.numpy(), pred_small=b.numpy(), iou_small=quiet(ref.mask_iou, a, b).numpy(),
                dice_small=quiet(ref.mask_dice, a, b).numpy())
    out = os.path.join(ROOT, "tests", "golden", "validate_metrics.npz")
    np.savez_compressed(out, **pack)
    print("wrote", out, {k: v.shape for k, v in pack.items()})


if __name__ == "__main__":
    main()
