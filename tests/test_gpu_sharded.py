"""GPU: the Z-sharded pass (all ranks emulated in one process on one GPU, collectives replaced by
copies) must be bit-identical to the unsharded pass — same labels, same numbering."""
import numpy as np
import pytest
import torch

import skoots_oracle as orc
from skoots_b200.synthetic import make_tube_volume

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equals_unsharded_on_tubes(world, transport, split):
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    shape = (96, 80, 256)
    tv = make_tube_volume(shape, 120, seed=3, device=DEV)
    # make z matter: stretch some objects along z so they cross slab faces
    tv.skeleton[40:43, 30:33, 20:240] = 1
    tv.vectors[2, 36:47, 26:37, :] = 0.5
    scale = (60, 60, 12)
    want = assemble_instances(tv.skeleton, tv.vectors, torch.tensor(scale), N=1)
    grp = LocalGroup(shape, world, DEV, scale=scale, transport=transport, split=split)
    grp.load_volume(tv.skeleton, tv.vectors)
    got = grp.step()
    assert torch.equal(got, want)
    for _ in range(3):  # buffers are reused across passes (the peer transport alternates between two copies)
        assert torch.equal(grp.step(), want)
    assert int(want.max()) > 3


def test_sharded_halo_only_blobs_and_columns():
    """components that (a) span every slab, (b) live entirely inside a neighbour's halo planes and are
    only reached by vectors from the other side of the face, (c) touch a face from one side only."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    X, Y, Z = 32, 40, 256
    mask = torch.zeros((X, Y, Z), dtype=torch.uint8, device=DEV)
    mask[5, 5, :] = 1                 # (a) one column through all 4 slabs
    mask[10:12, 10:12, 66:70] = 1     # (b) blob just above the face at z=64
    mask[20, 20, 60:64] = 1           # (c) ends exactly at the face
    mask[20, 22, 64:70] = 1           #     starts exactly at the face (different component)
    mask[25, 30, 120:136] = 1         # crosses the face at 128
    vec = torch.zeros((3, X, Y, Z), dtype=torch.float16, device=DEV)
    vec[2, 10:12, 10:12, 56:64] = 0.75   # voxels below the face point 9 planes up, into the blob
    vec[2, 20, 22, 64:70] = -0.5         # voxels above the face point 6 planes down
    vec[2, 25, 30, 100:120] = 1.0
    scale = (60, 60, 12)
    want = assemble_instances(mask, vec, torch.tensor(scale), N=1)
    ref = orc.postprocess(mask.cpu(), vec.cpu(), torch.tensor(scale), N=1)
    assert torch.equal(want.cpu(), ref)
    for world, transport in ((2, "peer"), (4, "peer"), (4, "nccl")):
        grp = LocalGroup((X, Y, Z), world, DEV, scale=scale, transport=transport)
        grp.load_volume(mask, vec)
        assert torch.equal(grp.step(), want), (world, transport)
        assert torch.equal(grp.step(), want), (world, transport)
    assert int(want[10, 10, 60]) == int(want[10, 10, 67]) > 0


def test_peer_pass_replays_as_one_cuda_graph():
    """the peer transport has no host step inside a pass: every rank's pass captures into a CUDA graph and
    replays bit-identically (the pass counter lives on the device, so the buffer parity advances per replay)."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    shape = (64, 64, 128)
    tv = make_tube_volume(shape, 40, seed=5, device=DEV)
    tv.skeleton[30:32, 30:32, 10:120] = 1
    want = assemble_instances(tv.skeleton, tv.vectors, torch.tensor((60, 60, 12)), N=1)
    grp = LocalGroup(shape, 2, DEV, transport="peer")
    grp.load_volume(tv.skeleton, tv.vectors)
    assert torch.equal(grp.step(), want)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for r in grp.ranks:
                r.phase_local()
            for r in grp.ranks:
                r.phase_ingest()
            for r in grp.ranks:
                r.phase_merge_and_gather()
    for _ in range(3):
        for r in grp.ranks:
            r.out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(torch.cat([r.out for r in grp.ranks], dim=2), want)
        assert all(int(r.meta[1]) == 0 for r in grp.ranks)


@pytest.mark.parametrize("density", [0.05, 0.3, 0.7])
def test_sharded_dense_random_masks_and_fields(density):
    """dense random masks and vector fields: thousands of runs per face (the emit / boundary kernels leave their
    shared-memory queues for the in-place fallbacks), many face pairs, dense tiles — still bit-identical to the
    unsharded pass and to the oracle, for both transports."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    g = torch.Generator().manual_seed(int(density * 100))
    shape = (48, 40, 256)
    mask = (torch.rand(shape, generator=g) < density).to(torch.uint8)
    vec = ((torch.rand((3,) + shape, generator=g) * 2 - 1) * (torch.rand((3,) + shape, generator=g) < 0.5)).to(torch.float16)
    scale = (5, 5, 12)
    ref = orc.postprocess(mask, vec, torch.tensor(scale), N=1)
    want = assemble_instances(mask.to(DEV), vec.to(DEV), torch.tensor(scale), N=1)
    assert torch.equal(want.cpu(), ref)
    for world, transport in ((2, "peer"), (4, "peer"), (3, "nccl")):
        if 256 // 64 < world:
            continue
        grp = LocalGroup(shape, world, DEV, scale=scale, transport=transport)
        grp.load_volume(mask.to(DEV), vec.to(DEV))
        for _ in range(2):
            assert torch.equal(grp.step(), want), (world, transport)
