"""GPU: the CUDA path (through the ctypes C-ABI) against the fixtures produced by the unmodified
reference.  Integer outputs bit-exact; embeddings bit-exact (the kernel restates every fp32
rounding of the reference)."""
import numpy as np
import pytest
import torch

import skoots_oracle as orc
from conftest import load_golden, unpack_mask
from skoots_b200.synthetic import make_tube_volume

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def test_library_loaded_is_in_tree():
    import skoots_b200._lib as L
    assert L.load().skb_version() == 201
    assert L.LIB_PATH.endswith("skoots_b200/libskoots_b200.so")


def test_known_answer_vec2embed():
    from skoots_b200.lib.vector_to_embedding import vector_to_embedding
    fx = load_golden("kat_vec2embed")
    out = vector_to_embedding(cu(fx["scale"]), cu(fx["vector"]), N=int(fx["N"]))
    assert out[0, :, 5, 5, 5].tolist() == [6.0, 6.0, 6.0]
    assert np.array_equal(out.cpu().numpy(), fx["out"])


@pytest.mark.parametrize("tag,dt", [("f16", torch.float16), ("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_vector_to_embedding_bit_exact(tag, dt):
    from skoots_b200.lib.vector_to_embedding import vector_to_embedding
    fx = load_golden(f"vec2embed_{tag}")
    vec = cu(fx["vector"]).to(dt)
    for key in [k for k in fx.files if k.startswith("out_")]:
        N = int(key.split("_")[1][1:])
        decay = int(key.split("_")[2][1:]) / 100.0
        got = vector_to_embedding(cu(fx["scale"]), vec, N=N, decay=decay)
        assert got.dtype == torch.float32
        assert np.array_equal(got.cpu().numpy(), fx[key]), key


def test_vector_to_embedding_2d_and_asserts():
    from skoots_b200.lib.vector_to_embedding import vector_to_embedding
    fx = load_golden("vec2embed_2d")
    got = vector_to_embedding(cu(fx["scale"]), cu(fx["vector"]))
    assert np.array_equal(got.cpu().numpy(), fx["out"])
    with pytest.raises(AssertionError):
        vector_to_embedding(cu(fx["scale"]), cu(fx["vector"]), N=2)


def test_index_skeleton_by_embed():
    from skoots_b200.lib.skeleton import index_skeleton_by_embed
    fx = load_golden("index_by_embed")
    got = index_skeleton_by_embed(cu(fx["labels"]), cu(fx["embed"]))
    assert got.dtype == torch.int32 and got.shape == fx["out"].shape
    assert np.array_equal(got.cpu().numpy(), fx["out"])


@pytest.mark.parametrize("name", ["flood_small", "flood_dense"])
def test_flood_fill_identical_numbering(name):
    from skoots_b200.lib.flood_fill import efficient_flood_fill
    fx = load_golden(name)
    vol = cu(unpack_mask(fx)).to(torch.int16)
    out = efficient_flood_fill(vol)
    assert out.data_ptr() == vol.data_ptr()  # in place, like the reference
    assert np.array_equal(out.cpu().numpy(), fx["out"])


def test_flood_fill_multicrop_partition():
    import skoots_oracle as orc
    from skoots_b200.lib.flood_fill import efficient_flood_fill
    fx = load_golden("flood_multicrop")
    out = efficient_flood_fill(cu(unpack_mask(fx)).to(torch.int16).unsqueeze(0))
    assert out.shape == fx["out"].shape
    assert np.array_equal(orc.canonical_relabel(out.cpu().numpy()), orc.canonical_relabel(fx["out"]))


def test_flood_fill_reference_crops_bit_identical():
    """row f3: with reference_crops=True the multi-crop behaviour of the reference (per-crop numbering, seam
    heuristic, last-member replacement) is reproduced bit for bit on the reference's own output."""
    from skoots_b200.lib.flood_fill import _flood_fill_reference_crops, efficient_flood_fill
    fx = load_golden("flood_multicrop")
    vol = cu(unpack_mask(fx)).to(torch.int16)
    out = efficient_flood_fill(vol, reference_crops=True)
    assert out.data_ptr() == vol.data_ptr() and np.array_equal(out.cpu().numpy(), fx["out"])
    # small crops on a small volume against the oracle's replay: shifted last crops, several seams per axis, an empty crop
    tv = make_tube_volume((70, 52, 40), 40, seed=13)
    mask = tv.skeleton.clone()
    mask[:30, :20, :16] = 0                      # one empty crop -> the reference restarts its numbering (B#7)
    want = orc.flood_fill_multicrop(mask.to(torch.int16), crop=(30, 20, 16)).numpy()
    got = _flood_fill_reference_crops(mask.to(torch.int16).to(DEV), crop=(30, 20, 16))
    assert np.array_equal(got.cpu().numpy(), want)


def test_assembly_against_reference_loop():
    from skoots_b200.pipeline import assemble_instances, gather_instances
    fx = load_golden("assembly")
    mask = cu(unpack_mask(fx, "skeleton"))
    vec = cu(fx["vectors"]).to(torch.float16)
    scale = torch.from_numpy(fx["scale"])
    labels = cu(fx["labels"])
    for key in [k for k in fx.files if k.startswith("cfg_")]:
        N, d100, cx, cy, cz, ox, oy, oz = (int(v) for v in fx[key])
        want = fx["inst_" + key[4:]]
        fused = assemble_instances(mask, vec, scale, N=N, decay=d100 / 100.0, crop=(cx, cy, cz), overlap=(ox, oy, oz),
                                   out_dtype=torch.int16)
        assert np.array_equal(fused.cpu().numpy(), want), key
        dense = gather_instances(vec, scale, labels, N=N, decay=d100 / 100.0, crop=(cx, cy, cz), overlap=(ox, oy, oz))
        assert np.array_equal(dense.cpu().numpy(), want.astype(np.int32)), key
    for N in (1, 4):
        got = assemble_instances(mask, vec, scale, N=N)
        assert got.dtype == torch.int32
        assert np.array_equal(got.cpu().numpy(), fx[f"inst_whole_N{N}"])


@pytest.mark.parametrize("name", ["morphology", "morphology_binary"])
def test_morphology(name):
    from skoots_b200.lib.morphology import binary_dilation, binary_dilation_2d, binary_erosion
    fx = load_golden(name)
    img = cu(fx["image"])
    assert np.array_equal(binary_dilation(img).cpu().numpy(), fx["dilation"])
    assert np.array_equal(binary_dilation_2d(img).cpu().numpy(), fx["dilation_2d"])
    ero = binary_erosion(img).cpu().numpy()
    assert ero.shape == fx["erosion"].shape and np.array_equal(ero, fx["erosion"])


def test_tile_epilogue():
    from skoots_b200.pipeline import tile_epilogue
    fx = load_golden("tile_epilogue")
    vol_v = torch.zeros(fx["vectors"].shape, dtype=torch.float16, device=DEV)
    vol_s = torch.zeros(fx["skeleton"].shape, dtype=torch.uint8, device=DEV)
    tile_epilogue(cu(fx["unet"]), vol_v, vol_s, tuple(int(v) for v in fx["origin"]), tuple(int(v) for v in fx["overlap"]))
    assert np.array_equal(vol_v.float().cpu().numpy(), fx["vectors"])
    assert np.array_equal(vol_s.cpu().numpy(), fx["skeleton"])


def test_baked_embed_to_prob():
    from skoots_b200.lib.embedding_to_prob import baked_embed_to_prob
    fx = load_golden("embed_prob")
    got = baked_embed_to_prob(cu(fx["embedding"]), cu(fx["baked"]), cu(fx["sigma"]))
    assert got.shape == fx["out"].shape
    np.testing.assert_allclose(got.cpu().numpy(), fx["out"], rtol=1e-5, atol=0)  # north star: 1e-5 relative in fp32


def _skeleton_dict(fx):
    out, at = {}, 0
    for k, n in zip(fx["ids"], fx["lens"]):
        out[int(k)] = cu(fx["points"][at:at + int(n)])
        at += int(n)
    return out


def test_bake_skeleton():
    from skoots_b200.lib.skeleton import bake_skeleton
    fx = load_golden("bake_skeleton")
    sk = _skeleton_dict(fx)
    mask = cu(fx["mask"])
    for tag, an in (("iso", (1.0, 1.0, 1.0)), ("aniso", (1.0, 1.0, 3.0))):
        got = bake_skeleton(mask, sk, anisotropy=an, average=False)
        assert np.array_equal(got.cpu().numpy(), fx[f"baked_{tag}"]), tag
        got = bake_skeleton(mask.unsqueeze(0), sk, anisotropy=an, average=True)
        np.testing.assert_allclose(got.cpu().numpy(), fx[f"baked_avg_{tag}"], rtol=1e-5, atol=1e-6)
    with pytest.raises(KeyError):
        bake_skeleton(mask, {k: v for k, v in list(sk.items())[1:]})
    assert bake_skeleton(mask, {-1: sk[next(iter(sk))]}).dtype == torch.float16


@pytest.mark.parametrize("name", ("blobs", "aniso", "missing", "crop"))
def test_bake_skeleton_triton_compat_against_reference_fixture(name):
    """a8 with the semantics of the reference's own GPU kernel (triton_compat=True): bit-equal to what the unmodified
    reference's Triton launch returned on a B200 (tests/golden/bake_triton.npz) — raw fp16 points, fp16 distances,
    and the averaged field; one call per sample and the batch call."""
    from test_oracle_golden import triton_case
    from skoots_b200.lib.skeleton import bake_skeleton, bake_skeletons_batch
    fx = load_golden("bake_triton")
    mask, sk, an = triton_case(fx, name)
    mask = mask.to(DEV)
    skd = {k: v.to(DEV).float() for k, v in sk.items()}
    raw, dist = bake_skeleton(mask, skd, an, average=False, return_distance=True, triton_compat=True)
    assert raw.dtype == torch.float16 and dist.dtype == torch.float16 and dist.shape == (1,) + tuple(mask.shape)
    assert np.array_equal(raw.cpu().numpy(), fx[f"{name}_raw"])
    ulp = np.abs(dist.cpu().numpy().view(np.int16).astype(np.int32) - fx[f"{name}_dist"].view(np.int16).astype(np.int32))
    assert ulp.max() <= 1
    avg = bake_skeleton(mask, skd, an, average=True, triton_compat="auto")   # a CUDA mask: the reference dispatches Triton
    assert avg.dtype == torch.float32
    np.testing.assert_allclose(avg.cpu().numpy(), fx[f"{name}_avg"], rtol=1e-5, atol=1e-6)
    both = bake_skeletons_batch(torch.stack([mask, mask.flip(0)]), [skd, skd], an, average=False, triton_compat=True)
    assert np.array_equal(both[0].cpu().numpy(), fx[f"{name}_raw"].astype(np.float32))
    # host tensors under "auto" keep the CPU semantics (and its KeyError for the id without a skeleton)
    if name == "missing":
        with pytest.raises(KeyError):
            bake_skeleton(mask.cpu(), sk, an, average=False, triton_compat="auto")
    else:
        host = bake_skeleton(mask.cpu(), sk, an, average=False, triton_compat="auto")
        want = orc.bake_skeleton(mask.cpu(), sk, an, average=False)
        assert host.dtype == torch.float32 and torch.equal(host, want)


def test_average_baked():
    from skoots_b200.lib.skeleton import average_baked_skeletons
    fx = load_golden("average_baked")
    np.testing.assert_allclose(average_baked_skeletons(cu(fx["baked"])).cpu().numpy(), fx["out"], rtol=1e-5, atol=1e-7)


def test_skeleton_to_mask():
    from skoots_b200.lib.skeleton import get_cached_disk_coords, skeleton_to_mask
    fx = load_golden("skeleton_to_mask")
    pts = fx["points"]
    sk = {1: cu(pts[:3]), 2: cu(pts[3:])}
    for r, f in ((7, 3), (9, 3), (2, 1)):
        assert np.array_equal(get_cached_disk_coords(DEV, r, f).cpu().numpy(), fx[f"offsets_r{r}_f{f}"])
        got = skeleton_to_mask(sk, (40, 36, 8), radius=r, flank_radius=f)
        assert np.array_equal(got.cpu().numpy(), fx[f"mask_r{r}_f{f}"])


def test_2d_mode_fused_gather_matches_reference_fixture():
    """a10 / BASELINE configs[4]: planar CCL + the fused 2-D gather against the reference-generated fixture, then
    ragged image sizes (no 16-byte alignment, plane not a multiple of 8) and 16-bit output against the oracle."""
    from skoots_b200.lib.flood_fill import label_components, write_dense
    from skoots_b200.pipeline import assemble_instances_2d, gather_instances_2d
    fx = load_golden("assembly_2d")
    masks = cu(unpack_mask(fx, "masks"))
    for dt in (torch.float32, torch.float16):
        vec = cu(fx["vectors"]).to(dt)
        want = fx["out"] if dt == torch.float32 else orc.postprocess_2d(masks.cpu(), vec.cpu(), torch.from_numpy(fx["scale"])).numpy()
        got = assemble_instances_2d(masks, vec, torch.from_numpy(fx["scale"]))
        assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want), dt
    sp = label_components(masks, planar=True, label_base=0)
    dense = torch.empty(masks.shape, dtype=torch.int32, device=DEV)
    write_dense(sp, dense)
    assert np.array_equal(gather_instances_2d(cu(fx["vectors"]), torch.from_numpy(fx["scale"]), dense).cpu().numpy(), fx["out"])
    g = torch.Generator().manual_seed(2)
    for shape in ((2, 37, 29), (5, 64, 128), (1, 8, 8)):
        m = (torch.rand(shape, generator=g) < 0.2).to(torch.uint8)
        v = ((torch.rand((shape[0], 2) + shape[1:], generator=g) * 2 - 1) * (torch.rand((shape[0], 2) + shape[1:], generator=g) < 0.6)).to(torch.float16)
        scale = torch.tensor((7.0, 5.0))
        want = orc.postprocess_2d(m, v, scale)
        got = assemble_instances_2d(m.to(DEV), v.to(DEV), scale, out_dtype=torch.int16)
        assert np.array_equal(got.cpu().numpy(), want.numpy().astype(np.int16)), shape


def test_elastic_deform_f4():
    """elastic_deform without the dense grids.  Parity bar: the displacement is a chain of fp32
    lerps that ATen itself evaluates differently on CPU and CUDA, so samples may differ where a coordinate lies within
    ~1e-4 of a rounding boundary: at most 0.1 % of the voxels may differ from the reference's output, and a skeleton
    coordinate (integer after truncation) by at most 1."""
    from skoots_b200.train.merged_transform import elastic_deform
    fx = load_golden("elastic")
    for tag in ("a", "b"):
        sk = {1: cu(fx[f"{tag}_sk1"]), 2: cu(fx[f"{tag}_sk2"])}
        ds = tuple(int(v) for v in fx[f"{tag}_ds"])
        img, mask, new = elastic_deform(cu(fx[f"{tag}_image"]), cu(fx[f"{tag}_mask"]), skeleton=sk, displacement_shape=ds,
                                        displacement_magnitude=tuple(float(v) for v in fx[f"{tag}_mag"]), noise=cu(fx[f"{tag}_noise"]))
        for got, key in ((img, "image"), (mask, "mask")):
            want = fx[f"{tag}_out_{key}"]
            assert got.shape == want.shape
            assert float((got.cpu().numpy() != want).mean()) <= 1e-3, (tag, key)
        for k in (1, 2):
            want = fx[f"{tag}_out_sk{k}"]
            assert new[k].dtype == torch.int64 and int(np.abs(new[k].cpu().numpy() - want).max()) <= 1, (tag, k)
            assert float((new[k].cpu().numpy() != want).mean()) <= 0.05
        assert np.array_equal(new[2].cpu().numpy()[2:4], fx[f"{tag}_sk2"][2:4])   # outside the volume: unchanged
    # a seeded call draws the field exactly like the reference does (torch.rand of the same shape on the same device)
    a = cu(fx["a_image"])
    torch.manual_seed(3)
    one = elastic_deform(a, skeleton={})[0]
    torch.manual_seed(3)
    two = elastic_deform(a, skeleton={}, noise=torch.rand((1, 3, 2, 6, 6), device=DEV))[0]
    assert torch.equal(one, two)
