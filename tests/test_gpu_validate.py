"""GPU: row f2 — renumber (fastremap.renumber, eval.py:304) and the validation metrics (validate/lib.py) against
the fixtures generated from the unmodified reference and against the oracle on larger seeded masks."""
import numpy as np
import pytest
import torch

import skoots_oracle as orc
from conftest import load_golden
from skoots_b200.synthetic import make_tube_volume

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_metrics_match_reference_fixtures():
    from skoots_b200.validate import accuracies_from_iou, mask_dice, mask_iou
    fx = load_golden("validate_metrics")
    for sfx in ("", "_small"):
        gt, pred = torch.from_numpy(fx["gt" + sfx]).to(DEV), torch.from_numpy(fx["pred" + sfx]).to(DEV)
        assert np.array_equal(mask_iou(gt, pred).cpu().numpy(), fx["iou" + sfx])      # bit-exact fp32
        assert np.array_equal(mask_dice(gt, pred).cpu().numpy(), fx["dice" + sfx])
    iou = torch.from_numpy(fx["iou"]).to(DEV)
    for thr in (0.1, 0.3, 0.5, 0.75):
        assert list(accuracies_from_iou(iou, thr)) == fx[f"acc_{int(thr * 100)}"].tolist()


@pytest.mark.parametrize("dtypes", [(torch.int32, torch.int32), (torch.int16, torch.int32), (torch.int32, torch.int16)])
def test_metrics_vs_oracle_on_tubes(dtypes):
    from skoots_b200.validate import accuracies_from_iou, mask_dice, mask_iou
    tv = make_tube_volume((96, 80, 33), 60, seed=4)   # ragged Z: the unaligned load path
    gt = tv.mask.to(dtypes[0])
    pred = torch.roll(tv.mask, shifts=(2, 0, 1), dims=(0, 1, 2)).to(dtypes[1])
    pred[pred == 5] = 6
    got = mask_iou(gt.to(DEV), pred.to(DEV))
    want = orc.mask_iou(gt.numpy(), pred.numpy())
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    assert torch.equal(mask_dice(gt.to(DEV), pred.to(DEV)).cpu(), orc.mask_dice(gt.numpy(), pred.numpy()))
    for thr in (0.05, 0.4):
        assert accuracies_from_iou(got, thr) == orc.accuracies_from_iou(want, thr)


def test_metrics_edge_cases():
    import skoots_b200._lib as L
    from skoots_b200.validate import accuracies_from_iou, mask_dice, mask_iou
    z = torch.zeros((4, 5, 6), dtype=torch.int32, device=DEV)
    one = z.clone()
    one[1:3, 1:3, 1:3] = 4
    assert tuple(mask_iou(z, z).shape) == (0, 0)
    assert tuple(mask_iou(one, z).shape) == (1, 0) and tuple(mask_iou(z, one).shape) == (0, 1)
    neg = one.clone()
    neg[0, 0, 0] = -3                                  # `unique > 0`: negatives are background to the reference
    assert torch.equal(mask_iou(neg, one).cpu(), orc.mask_iou(neg.cpu().numpy(), one.cpu().numpy()))
    assert float(mask_iou(one, one)[0, 0]) == 1.0
    with pytest.raises(AssertionError):
        mask_dice(one, one)                            # the reference's assert (validate/lib.py:266-268)
    with pytest.raises(IndexError):
        accuracies_from_iou(mask_iou(one, z))
    host = mask_iou(one.cpu(), one.cpu())              # host tensors are staged through the GPU and come back on the host
    assert not host.is_cuda and float(host[0, 0]) == 1.0
    with pytest.raises(AssertionError):
        mask_iou(one.cpu(), one)                       # a host / device mix: the reference's own assert (validate/lib.py:199)
    with pytest.raises(AssertionError):
        mask_iou(one, one[:2])


@pytest.mark.parametrize("dt", [torch.int16, torch.int32])
def test_renumber_matches_first_occurrence_order(dt):
    from skoots_b200.validate import renumber
    g = torch.Generator().manual_seed(2)
    vol = (torch.randint(0, 40, (37, 21, 19), generator=g) * 13).to(dt)
    vol[vol == 13 * 7] = 0
    want, remap = orc.renumber(vol.numpy())
    got, table = renumber(vol.to(DEV))
    assert got.dtype == dt and np.array_equal(got.cpu().numpy(), want)
    table = table.cpu().numpy()
    for old, new in remap.items():
        assert table[old] == new
    assert int((table > 0).sum()) == len(remap) - 1
    # in place, on the output of the assembly (labels 3..N+2 with gaps where a component caught no voxel)
    from skoots_b200.pipeline import assemble_instances
    tv = make_tube_volume((64, 64, 64), 30, seed=9, device=DEV)
    inst = assemble_instances(tv.skeleton, tv.vectors, torch.tensor((60, 60, 12)), N=1)
    want, _ = orc.renumber(inst.cpu().numpy())
    out, _ = renumber(inst, in_place=True)
    assert out.data_ptr() == inst.data_ptr() and np.array_equal(inst.cpu().numpy(), want)
    assert int(inst.max()) == len(np.unique(want)) - 1


def test_renumber_large_volume_properties():
    """size-independent properties at 512x512x256: idempotent, a bijection on the labels present, background kept."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.validate import label_max, renumber
    tv = make_tube_volume((512, 512, 256), 500, seed=1, device=DEV, want_mask=False, want_skeleton_dict=False)
    inst = assemble_instances(tv.skeleton, tv.vectors, torch.tensor((60, 60, 12)), N=1)
    out, remap = renumber(inst)
    assert torch.equal(out == 0, inst == 0)
    n = label_max(out)
    assert n == int((remap > 0).sum()) == int(torch.unique(inst).numel()) - 1
    again, remap2 = renumber(out)
    assert torch.equal(again, out) and torch.equal(remap2[1:], torch.arange(1, n + 1, device=DEV, dtype=torch.int32))
    assert torch.equal(remap[inst.long()], out)          # out is exactly the table applied to the input
    # first appearance order: label k+1 first appears after label k
    flat = out.reshape(-1)
    first = torch.full((n + 1,), flat.numel(), dtype=torch.int64, device=DEV)
    first.scatter_reduce_(0, flat.long(), torch.arange(flat.numel(), device=DEV), reduce="amin")
    assert bool((first[2:] > first[1:-1]).all())
