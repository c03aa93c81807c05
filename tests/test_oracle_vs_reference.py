"""CPU, build container only: the oracle restatement against the UNMODIFIED reference functions run
live (skipped where /root/reference does not exist, e.g. the GPU box — the committed fixtures cover
that case)."""
import contextlib
import io

import numpy as np
import pytest
import torch

import ref_shim
import skoots_oracle as orc
from skoots_b200.synthetic import make_tube_volume

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    ref_shim.install()
    import skoots.lib.embedding_to_prob as e2p
    import skoots.lib.flood_fill as ff
    import skoots.lib.morphology as morph
    import skoots.lib.skeleton as skel
    import skoots.lib.vector_to_embedding as v2e

    class R:
        pass
    r = R()
    r.v2e, r.index, r.flood = v2e.vector_to_embedding, skel.index_skeleton_by_embed, ff.efficient_flood_fill
    r.dil, r.dil2d, r.ero = morph.binary_dilation, morph.binary_dilation_2d, morph.binary_erosion
    r.prob, r.bake, r.s2m = e2p.baked_embed_to_prob, skel.bake_skeleton, skel.skeleton_to_mask
    return r


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("seed", range(6))
def test_vec2embed_random_fields(ref, seed):
    g = torch.Generator().manual_seed(seed)
    shape = [int(v) for v in torch.randint(3, 20, (3,), generator=g)]
    vec = ((torch.rand([1, 3] + shape, generator=g) * 2 - 1) * 2.5).to([torch.float16, torch.bfloat16, torch.float32][seed % 3])
    scale = torch.randint(1, 12, (3,), generator=g)
    for N, decay in ((1, 1.0), (3, 1.0), (7, 0.9)):
        assert torch.equal(orc.vector_to_embedding(scale, vec, N, decay), ref.v2e(scale, vec, N, decay)), (shape, N, decay)


def test_vec2embed_batch_quirk(ref):
    """B > 1 with N > 1: the reference's `take` runs over the flattened (B,1,X,Y,Z) slice with per-volume indices, so
    the hops of EVERY batch element read the vectors of element 0 (vector_to_embedding.py:130)."""
    g = torch.Generator().manual_seed(77)
    vec = ((torch.rand((3, 3, 9, 8, 7), generator=g) * 2 - 1) * 1.5).to(torch.float16)
    scale = torch.tensor((4, 3, 2))
    for N, decay in ((1, 1.0), (3, 1.0), (5, 0.9)):
        assert torch.equal(orc.vector_to_embedding(scale, vec, N, decay), ref.v2e(scale, vec, N, decay)), (N, decay)


@pytest.mark.parametrize("seed", range(4))
def test_flood_fill_random_masks(ref, seed):
    g = torch.Generator().manual_seed(100 + seed)
    shape = [int(v) for v in torch.randint(4, 40, (3,), generator=g)]
    mask = (torch.rand(shape, generator=g) < [0.05, 0.3, 0.5, 0.8][seed]).to(torch.int16)
    assert torch.equal(orc.flood_fill_exact(mask.clone()), quiet(ref.flood, mask.clone()))


def test_c1_postprocess_equals_reference_functions(ref):
    tv = make_tube_volume((128, 128, 32), 20, seed=0)
    scale = torch.tensor((60, 60, 12))
    labels = quiet(ref.flood, tv.skeleton.to(torch.int16).clone())
    for N in (1, 10):
        want = ref.index(labels[None, None], ref.v2e(scale, tv.vectors[None], N=N))[0, 0]
        assert torch.equal(orc.postprocess(tv.skeleton, tv.vectors, scale, N=N), want)


def test_training_ops(ref):
    tv = make_tube_volume((60, 50, 12), 6, seed=4)
    present = {int(k): tv.skeletons[int(k)] for k in torch.unique(tv.mask).tolist() if k != 0}
    for an in ((1.0, 1.0, 1.0), (1.0, 1.0, 3.0)):
        assert torch.equal(orc.bake_skeleton(tv.mask, present, an, average=False), ref.bake(tv.mask, present, an, average=False))
    assert torch.equal(orc.skeleton_to_mask(present, (60, 50, 12), 9, 3), ref.s2m(present, (60, 50, 12), radius=9, flank_radius=3))
    E, S = torch.rand((2, 3, 8, 7, 6)) * 40, torch.rand((2, 3, 8, 7, 6)) * 40
    sig = torch.tensor((20.0, 20.0, 20.0))
    np.testing.assert_allclose(orc.baked_embed_to_prob(E, S, sig).numpy(), ref.prob(E, S, sig).numpy(), rtol=1e-6)
    img = torch.randn((1, 2, 9, 8, 7))
    assert torch.equal(orc.binary_dilation(img), ref.dil(img))
    assert torch.equal(orc.binary_dilation_2d(img), ref.dil2d(img))
    assert torch.equal(orc.binary_erosion(img), ref.ero(img))


def test_validation_metrics_live(ref):
    """row f2: the oracle's contingency-table restatement against skoots.validate.lib run live."""
    import skoots.validate.lib as vl
    g = torch.Generator().manual_seed(33)
    gt = (torch.randint(0, 9, (14, 12, 10), generator=g) * 3).to(torch.int32)
    pred = torch.roll(gt, shifts=(1, 1, 0), dims=(0, 1, 2)).clone()
    pred[pred == 6] = 21
    pred[:3] = 0
    iou = quiet(vl.mask_iou, gt, pred)
    assert torch.equal(orc.mask_iou(gt.numpy(), pred.numpy()), iou)
    assert torch.equal(orc.mask_dice(gt.numpy(), pred.numpy()), quiet(vl.mask_dice, gt, pred))
    for thr in (0.05, 0.2, 0.6):
        assert orc.accuracies_from_iou(iou, thr) == tuple(vl.accuracies_from_iou(iou, thr))
