"""GPU: the stream/resolve split of the N = 1 gather (run next to the labelling chain on two streams) must
write exactly what the fused gather after the labelling writes — and what the CPU oracle computes."""
import pytest
import torch

import skoots_oracle as orc
from skoots_b200.synthetic import make_tube_volume

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALE = (60, 60, 12)


def _both(mask, vec, scale=SCALE, out_dtype=torch.int32):
    from skoots_b200.pipeline import assemble_instances
    a = assemble_instances(mask, vec, torch.tensor(scale), N=1, out_dtype=out_dtype, fused=False)
    b = assemble_instances(mask, vec, torch.tensor(scale), N=1, out_dtype=out_dtype, fused=True)
    return a, b


@pytest.mark.parametrize("shape,tubes", [((64, 64, 64), 30), ((96, 80, 256), 120), ((40, 52, 128), 40), ((8, 8, 64), 2)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
def test_split_equals_fused_and_oracle_on_tubes(shape, tubes, dt):
    tv = make_tube_volume(shape, tubes, seed=11, device=DEV)
    vec = tv.vectors.to(dt)
    a, b = _both(tv.skeleton, vec)
    assert torch.equal(a, b)
    want = orc.postprocess(tv.skeleton.cpu(), vec.cpu(), torch.tensor(SCALE), N=1)
    assert torch.equal(a.cpu(), want)
    assert shape == (8, 8, 64) or int(a.max()) > 2


@pytest.mark.parametrize("out_dtype", [torch.int16, torch.int32])
def test_split_random_fields_and_clamping(out_dtype):
    """dense random vectors: most targets leave the volume and are clamped; every 8-voxel group is flagged."""
    g = torch.Generator().manual_seed(3)
    shape = (24, 32, 128)
    mask = (torch.rand(shape, generator=g) < 0.2).to(torch.uint8).to(DEV)
    vec = ((torch.rand((3,) + shape, generator=g) * 2 - 1) * (torch.rand((3,) + shape, generator=g) < 0.5)).to(torch.float16).to(DEV)
    a, b = _both(mask, vec, out_dtype=out_dtype)
    assert torch.equal(a, b)
    want = orc.postprocess(mask.cpu(), vec.cpu(), torch.tensor(SCALE), N=1, out_dtype=out_dtype)
    assert torch.equal(a.cpu(), want)


def test_split_zero_vectors_label_themselves():
    """no vector anywhere: only skeleton voxels get a label (their own); -0.0 counts as zero."""
    shape = (16, 16, 128)
    mask = torch.zeros(shape, dtype=torch.uint8, device=DEV)
    mask[3, 4, 10:90] = 1
    mask[8:10, 8:10, 64:70] = 1
    vec = torch.zeros((3,) + shape, dtype=torch.float16, device=DEV)
    vec[1, :, :, ::2] = -0.0
    a, b = _both(mask, vec)
    assert torch.equal(a, b)
    assert torch.equal(a > 0, mask > 0)
    assert sorted(torch.unique(a).tolist()) == [0, 3, 4]


def test_split_empty_and_full_masks():
    shape = (8, 16, 64)
    vec = torch.full((3,) + shape, 0.01, dtype=torch.float16, device=DEV)
    for fill in (0, 1):
        mask = torch.full(shape, fill, dtype=torch.uint8, device=DEV)
        a, b = _both(mask, vec)
        assert torch.equal(a, b)
        assert int(a.max()) == (3 if fill else 0)


def test_split_reuses_buffers_and_survives_capacity_overflow():
    from skoots_b200.lib.flood_fill import new_sparse
    from skoots_b200.pipeline import assemble_instances, assemble_split
    rng = torch.Generator().manual_seed(8)
    shape = (32, 32, 64)
    mask = (torch.rand(shape, generator=rng) < 0.5).to(torch.uint8).to(DEV)
    vec = ((torch.rand((3,) + shape, generator=rng) - 0.5) * 0.1).to(torch.float16).to(DEV)
    want = assemble_instances(mask, vec, torch.tensor(SCALE), N=1, fused=True)
    # explicit buffers, reused over passes on the same workspace
    sparse = new_sparse(shape, torch.device(DEV))
    out = torch.empty(shape, dtype=torch.int32, device=DEV)
    flags = torch.empty(mask.numel() // 256, dtype=torch.int32, device=DEV)
    for _ in range(3):
        out.fill_(-7)
        assemble_split(mask, vec, SCALE, sparse, out, group_flags=flags)
        assert torch.equal(out, want)
    # a tiny capacity overflows in the split pass; assemble_instances must notice and redo it
    tiny = new_sparse(shape, torch.device(DEV), capacity=16)
    assemble_split(mask, vec, SCALE, tiny, out, group_flags=flags)
    assert int(tiny.status.item()) & 1


def test_split_rejects_ineligible_calls():
    import skoots_b200._lib as L
    from skoots_b200.pipeline import assemble_instances
    mask = torch.zeros((8, 8, 60), dtype=torch.uint8, device=DEV)
    vec = torch.zeros((3, 8, 8, 60), dtype=torch.float16, device=DEV)
    with pytest.raises(L.SkootsB200Error):
        assemble_instances(mask, vec, torch.tensor(SCALE), N=1, fused=False)   # Z % 64 != 0
    assert assemble_instances(mask, vec, torch.tensor(SCALE), N=1).shape == (8, 8, 60)  # default: fused
    mask = torch.zeros((8, 8, 64), dtype=torch.uint8, device=DEV)
    vec = torch.zeros((3, 8, 8, 64), dtype=torch.float16, device=DEV)
    with pytest.raises(L.SkootsB200Error):
        assemble_instances(mask, vec, torch.tensor(SCALE), N=2, fused=False)
