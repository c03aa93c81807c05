"""GPU: the Z-sharded pass (all ranks emulated in one process on one GPU, collectives replaced by
copies) must be bit-identical to the unsharded pass — same labels, same numbering."""
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


def test_vectors_beyond_the_halo_are_detected_not_mislabelled():
    """nothing bounds the UNet's vectors; a target farther than the label halo from the slab used to be
    answered from the wrong halo word.  It is now reported, and a deeper halo gives the unsharded result."""
    import skoots_b200._lib as L
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    X, Y, Z = 32, 32, 256
    mask = torch.zeros((X, Y, Z), dtype=torch.uint8, device=DEV)
    mask[8:10, 8:10, 90:100] = 1
    mask[20, 20, 30:40] = 1
    vec = torch.zeros((3, X, Y, Z), dtype=torch.float16, device=DEV)
    vec[2, 8:10, 8:10, 50:64] = 3.0      # 36 planes up, across the face at z = 64, into the first blob
    vec[2, 20, 20, 64:80] = -2.5          # 30 planes down, across the same face, into the second
    scale = (60, 60, 12)
    want = assemble_instances(mask, vec, torch.tensor(scale), N=1)
    assert int(want[8, 8, 60]) == int(want[8, 8, 95]) > 0 and int(want[20, 20, 66]) == int(want[20, 20, 35]) > 0
    for transport in ("peer", "nccl"):
        grp = LocalGroup((X, Y, Z), 4, DEV, scale=scale, transport=transport)   # default halo = ceil(scale_z) = 12 planes
        grp.load_volume(mask, vec)
        with pytest.raises(L.SkootsB200Error, match="beyond this rank's slab"):
            grp.step()
        deep = LocalGroup((X, Y, Z), 4, DEV, scale=scale, transport=transport, halo=64)
        deep.load_volume(mask, vec)
        assert torch.equal(deep.step(), want)
        assert torch.equal(deep.step(), want)


def test_slab_gather_by_ranges_equals_whole_slab():
    """skb_assemble_slab_ex over X-ranges (what the pipelined host pass issues) writes what one whole-slab launch writes."""
    from skoots_b200.sharded import LocalGroup
    shape = (64, 48, 128)
    tv = make_tube_volume(shape, 60, seed=9, device=DEV)
    grp = LocalGroup(shape, 2, DEV, transport="nccl")
    grp.load_volume(tv.skeleton, tv.vectors)
    want = grp.step()
    for r in grp.ranks:
        whole = r.out.clone()
        r.out.fill_(-7)
        plane = shape[1] * r.Zl
        for x0, x1 in ((0, 8), (8, 40), (40, 64)):
            r.gather((x0 * plane, (x1 - x0) * plane))
        assert torch.equal(r.out, whole)
    assert int(want.max()) > 3


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("world,crop,ov,N,decay", [(2, (48, 40, 50), (4, 4, 5), 10, 1.0), (4, (48, 40, 50), (4, 4, 5), 4, 0.95),
                                                   (4, (40, 40, 16), (5, 5, 2), 6, 1.0), (2, None, (0, 0, 0), 3, 1.0)])
def test_sharded_multi_hop_walks_equal_unsharded(world, crop, ov, N, decay, transport):
    """eval()'s configuration (crop grid, N = 10) Z-sharded.  Hops that leave the slab inside their crop read
    the neighbour's vector planes from the vector halo; the result must equal the unsharded crop-grid pass bit for bit."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import LocalGroup
    shape = (96, 80, 256) if crop is not None else (48, 40, 128)
    tv = make_tube_volume(shape, 150, seed=8, device=DEV, scale=(20.0, 20.0, 8.0))
    tv.skeleton[40:43, 30:33, 20:shape[2] - 16] = 1
    tv.vectors[2, 30:50, 26:37, :] = 0.6     # walks that run along z, across crop seams and slab faces
    scale = (20, 20, 8)
    want = assemble_instances(tv.skeleton, tv.vectors, torch.tensor(scale), N=N, decay=decay, crop=crop, overlap=ov, out_dtype=torch.int16)
    if crop is None and world > 1:
        # one crop = the whole volume: a hop may land anywhere, which no neighbour halo can cover
        with pytest.raises(ValueError):
            LocalGroup(shape, world, DEV, scale=scale, transport=transport, hops=N, out_dtype=torch.int16)
        return
    grp = LocalGroup(shape, world, DEV, scale=scale, transport=transport, hops=N, decay=decay, crop=crop, overlap=ov, out_dtype=torch.int16)
    grp.load_volume(tv.skeleton, tv.vectors)
    for _ in range(2):
        got = grp.step()
        assert got.dtype == torch.int16 and torch.equal(got, want), (world, crop, N)
    assert int((want > 0).sum()) > 1000
