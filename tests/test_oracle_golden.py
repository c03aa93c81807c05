"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  These are the pins that make the oracle trustworthy."""
import numpy as np
import pytest
import torch

import skoots_oracle as orc
from conftest import load_golden, unpack_mask


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x if dtype is None else x.to(dtype)


def test_known_answers_from_reference_main_blocks():
    fx = load_golden("kat_vec2embed")
    out = orc.vector_to_embedding(t(fx["scale"]), t(fx["vector"]), N=int(fx["N"]))
    assert out[0, :, 5, 5, 5].tolist() == [6.0, 6.0, 6.0]  # vector_to_embedding.py:221-232
    assert np.array_equal(out.numpy(), fx["out"])


@pytest.mark.parametrize("tag,dt", [("f16", torch.float16), ("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_vector_to_embedding_bit_exact(tag, dt):
    fx = load_golden(f"vec2embed_{tag}")
    vec = t(fx["vector"]).to(dt)
    for key in [k for k in fx.files if k.startswith("out_")]:
        N = int(key.split("_")[1][1:])
        decay = int(key.split("_")[2][1:]) / 100.0
        got = orc.vector_to_embedding(t(fx["scale"]), vec, N=N, decay=decay).numpy()
        assert np.array_equal(got, fx[key]), key


def test_vector_to_embedding_2d():
    fx = load_golden("vec2embed_2d")
    got = orc.vector_to_embedding(t(fx["scale"]), t(fx["vector"])).numpy()
    assert np.array_equal(got, fx["out"])


def test_index_skeleton_by_embed():
    fx = load_golden("index_by_embed")
    got = orc.index_skeleton_by_embed(t(fx["labels"]), t(fx["embed"]))
    assert got.dtype == torch.int32
    assert np.array_equal(got.numpy(), fx["out"])


@pytest.mark.parametrize("name", ["flood_small", "flood_dense"])
def test_flood_fill_single_crop(name):
    fx = load_golden(name)
    mask = unpack_mask(fx)
    got = orc.flood_fill_exact(t(mask).to(torch.int16)).numpy()
    assert got.dtype == np.int16
    assert np.array_equal(got, fx["out"])  # identical numbering: scipy order + 2
    assert got[got > 0].min() == 3


def test_flood_fill_multicrop_replay_and_exact_agree():
    fx = load_golden("flood_multicrop")
    mask = unpack_mask(fx)
    replay = orc.flood_fill_multicrop(t(mask).to(torch.int16)).numpy()
    assert np.array_equal(replay, fx["out"])  # bug-compatible replay is bit-identical
    exact = orc.flood_fill_exact(t(mask).to(torch.int16)).numpy()
    # on this input the seam heuristic makes no false merge: same partition up to relabelling
    assert np.array_equal(orc.canonical_relabel(exact), orc.canonical_relabel(fx["out"]))


@pytest.mark.parametrize("name", ["morphology", "morphology_binary"])
def test_morphology(name):
    fx = load_golden(name)
    img = t(fx["image"])
    assert np.array_equal(orc.binary_dilation(img).numpy(), fx["dilation"])
    assert np.array_equal(orc.binary_dilation_2d(img).numpy(), fx["dilation_2d"])
    ero = orc.binary_erosion(img).numpy()
    assert ero.shape == fx["erosion"].shape
    assert np.array_equal(ero, fx["erosion"])


def test_tile_epilogue():
    fx = load_golden("tile_epilogue")
    vol_v = torch.zeros(fx["vectors"].shape, dtype=torch.float16)
    vol_s = torch.zeros(fx["skeleton"].shape, dtype=torch.uint8)
    orc.tile_epilogue(t(fx["unet"]), vol_v, vol_s, tuple(fx["origin"]), tuple(int(v) for v in fx["overlap"]))
    assert np.array_equal(vol_v.float().numpy(), fx["vectors"])
    assert np.array_equal(vol_s.numpy(), fx["skeleton"])


def test_assembly_crop_grid_and_whole_volume():
    fx = load_golden("assembly")
    labels = t(fx["labels"])
    vec = t(fx["vectors"]).to(torch.float16)
    scale = t(fx["scale"])
    assert np.array_equal(orc.flood_fill_exact(t(unpack_mask(fx, "skeleton")).to(torch.int16)).numpy(), fx["labels"])
    for key in [k for k in fx.files if k.startswith("cfg_")]:
        N, d100, cx, cy, cz, ox, oy, oz = (int(v) for v in fx[key])
        got = orc.assemble_instances(labels, vec, scale, N=N, decay=d100 / 100.0, crop=(cx, cy, cz),
                                     overlap=(ox, oy, oz))
        assert np.array_equal(got.numpy(), fx["inst_" + key[4:]]), key
    for N in (1, 4):
        emb = orc.vector_to_embedding(scale, vec[None], N=N)
        got = orc.index_skeleton_by_embed(labels[None, None], emb)[0, 0].numpy()
        assert np.array_equal(got, fx[f"inst_whole_N{N}"])
    whole = orc.postprocess(t(unpack_mask(fx, "skeleton")), vec, scale, N=1).numpy()
    assert np.array_equal(whole, fx["inst_whole_N1"])


def test_baked_embed_to_prob():
    fx = load_golden("embed_prob")
    got = orc.baked_embed_to_prob(t(fx["embedding"]), t(fx["baked"]), t(fx["sigma"])).numpy()
    np.testing.assert_allclose(got, fx["out"], rtol=1e-6, atol=0)


def _skeleton_dict(fx):
    out, at = {}, 0
    for k, n in zip(fx["ids"], fx["lens"]):
        out[int(k)] = t(fx["points"][at:at + int(n)])
        at += int(n)
    return out


def test_bake_skeleton_cpu_semantics():
    fx = load_golden("bake_skeleton")
    sk = _skeleton_dict(fx)
    mask = t(fx["mask"])
    for tag, an in (("iso", (1.0, 1.0, 1.0)), ("aniso", (1.0, 1.0, 3.0))):
        assert np.array_equal(orc.bake_skeleton(mask, sk, an, average=False).numpy(), fx[f"baked_{tag}"])
        np.testing.assert_allclose(orc.bake_skeleton(mask, sk, an, average=True).numpy(), fx[f"baked_avg_{tag}"],
                                   rtol=1e-6, atol=1e-6)


def test_average_baked():
    fx = load_golden("average_baked")
    np.testing.assert_allclose(orc.average_baked_skeletons(t(fx["baked"])).numpy(), fx["out"], rtol=1e-6, atol=1e-7)


def test_skeleton_to_mask_and_offsets():
    fx = load_golden("skeleton_to_mask")
    pts = fx["points"]
    sk = {1: t(pts[:3]), 2: t(pts[3:])}
    for r, f in ((7, 3), (9, 3), (2, 1)):
        off = orc.disk_stamp_offsets(r, f)
        assert np.array_equal(off.T, fx[f"offsets_r{r}_f{f}"])
        got = orc.skeleton_to_mask(sk, (40, 36, 8), radius=r, flank_radius=f).numpy()
        assert np.array_equal(got, fx[f"mask_r{r}_f{f}"])
    assert orc.disk_stamp_offsets(7, 3).shape[0] == 207


def test_canonical_relabel():
    a = np.array([[0, 7, 7], [3, 0, 7], [3, 9, 0]])
    assert orc.canonical_relabel(a).tolist() == [[0, 1, 1], [2, 0, 1], [2, 3, 0]]
    assert orc.canonical_relabel(np.zeros((2, 2), dtype=np.int16)).sum() == 0


def test_crop_origins_match_reference_quirks():
    # SURVEY B#16: X=2048 -> origins 0,400,800,1200,1548,1548 ; shifted last crop emitted twice
    assert orc.crop_origins(2048, 500, 50) == [0, 400, 800, 1200, 1548, 1548]
    assert orc.crop_origins(512, 50, 5) == [0, 40, 80, 120, 160, 200, 240, 280, 320, 360, 400, 440, 462]
    with pytest.raises(ValueError):
        orc.crop_origins(8, 8, 50)


# ---- f2: renumber + validation metrics ---------------------------------------------------------------
def test_validate_metrics_match_reference_outputs():
    """the oracle's contingency-table restatement of mask_iou / mask_dice / accuracies_from_iou against the
    outputs of the unmodified reference functions (oracle/gen_golden_validate.py)."""
    fx = load_golden("validate_metrics")
    for sfx in ("", "_small"):
        gt, pred = fx["gt" + sfx], fx["pred" + sfx]
        assert np.array_equal(orc.mask_iou(gt, pred).numpy(), fx["iou" + sfx])      # bit-exact fp32
        assert np.array_equal(orc.mask_dice(gt, pred).numpy(), fx["dice" + sfx])
    iou = torch.from_numpy(fx["iou"])
    for thr in (0.1, 0.3, 0.5, 0.75):
        assert list(orc.accuracies_from_iou(iou, thr)) == fx[f"acc_{int(thr * 100)}"].tolist()


def test_renumber_restatement():
    a = np.array([[0, 7, 7], [3, 0, 7], [3, 9, 0]], dtype=np.int16)
    out, remap = orc.renumber(a)
    assert out.tolist() == [[0, 1, 1], [2, 0, 1], [2, 3, 0]] and out.dtype == np.int16
    assert remap == {0: 0, 7: 1, 3: 2, 9: 3}
    with pytest.raises(AssertionError):
        orc.mask_dice(a, a)  # identical objects: the reference's own assert fires (validate/lib.py:266-268)


def test_2d_mode_against_reference_fixture():
    """a10: per-slice scipy label + the reference's 2-D embedding + its gather with Z = 1 (tests/golden/assembly_2d.npz)."""
    fx = load_golden("assembly_2d")
    masks = torch.from_numpy(unpack_mask(fx, "masks"))
    got = orc.postprocess_2d(masks, torch.from_numpy(fx["vectors"]), torch.from_numpy(fx["scale"]))
    assert np.array_equal(got.numpy(), fx["out"])


def test_elastic_deform_against_reference_fixture():
    """f4: the oracle's elastic_deform against tests/golden/elastic.npz — volumes produced by the unmodified reference
    (pinned); skeleton points restated (UNPINNED: the reference's skeleton loop cannot run under this torch, see
    oracle/gen_golden_elastic.py)."""
    fx = load_golden("elastic")
    for tag in ("a", "b"):
        sk = {1: torch.from_numpy(fx[f"{tag}_sk1"]), 2: torch.from_numpy(fx[f"{tag}_sk2"])}
        img, mask, new = orc.elastic_deform(torch.from_numpy(fx[f"{tag}_noise"]), torch.from_numpy(fx[f"{tag}_image"]),
                                            torch.from_numpy(fx[f"{tag}_mask"]), skeleton=sk,
                                            displacement_magnitude=tuple(float(v) for v in fx[f"{tag}_mag"]))
        assert np.array_equal(img.numpy(), fx[f"{tag}_out_image"]) and np.array_equal(mask.numpy(), fx[f"{tag}_out_mask"])
        assert np.array_equal(new[1].numpy(), fx[f"{tag}_out_sk1"]) and np.array_equal(new[2].numpy(), fx[f"{tag}_out_sk2"])
        assert np.array_equal(new[2].numpy()[2:4], fx[f"{tag}_sk2"][2:4])   # points outside the volume are left alone


TRITON_CASES = ("blobs", "aniso", "missing", "crop")


def triton_case(fx, name):
    sk, at = {}, 0
    for k, n in zip(fx[f"{name}_ids"].tolist(), fx[f"{name}_lens"].tolist()):
        sk[int(k)] = torch.from_numpy(fx[f"{name}_pts"][at:at + n].copy())
        at += n
    return torch.from_numpy(fx[f"{name}_mask"]), sk, tuple(float(v) for v in fx[f"{name}_anisotropy"])


@pytest.mark.parametrize("name", TRITON_CASES)
def test_bake_skeleton_triton_semantics_against_reference_fixture(name):
    """a8, the GPU half of the reference: the restatement of what its Triton kernel returns (fp16, anisotropy on the
    squared differences, phantom origin lanes, per-axis maximum on ties, zeros for an id without a skeleton) against
    tests/golden/bake_triton.npz — outputs of the unmodified reference on a B200 (oracle/gen_golden_triton.py)."""
    fx = load_golden("bake_triton")
    mask, sk, an = triton_case(fx, name)
    raw, dist = orc.bake_skeleton_triton(mask, sk, an, average=False)
    assert raw.dtype == torch.float16 and np.array_equal(raw.numpy(), fx[f"{name}_raw"])
    ulp = np.abs(dist.numpy().view(np.int16).astype(np.int32) - fx[f"{name}_dist"].view(np.int16).astype(np.int32))
    assert ulp.max() <= 1                                    # tl.sqrt is the approximate square root
    avg, _ = orc.bake_skeleton_triton(mask, sk, an, average=True)
    np.testing.assert_allclose(avg.numpy(), fx[f"{name}_avg"], rtol=1e-5, atol=1e-6)
    # and the fixture really exercises what sets the Triton kernel apart from the CPU path
    cpu = orc.bake_skeleton(mask, {**{int(k): torch.zeros((1, 3)) for k in np.unique(mask.numpy()) if k}, **sk}, an, average=False)
    assert ((raw.float() != cpu).any(0) & (mask != 0)).sum() > 100


def test_bake_triton_fixture_inputs_come_from_the_committed_generator():
    """the inputs stored in tests/golden/bake_triton.npz are exactly what oracle/gen_golden_triton.py's seeded cases()
    produce (the outputs next to them were computed by the reference on a B200 from those)."""
    import gen_golden_triton as gen
    fx = load_golden("bake_triton")
    made = gen.cases()
    assert tuple(made) == TRITON_CASES
    for name, (mask, sk, an) in made.items():
        ids, lens, pts = gen.pack_skeletons(sk)
        assert np.array_equal(mask, fx[f"{name}_mask"]) and np.array_equal(ids, fx[f"{name}_ids"])
        assert np.array_equal(lens, fx[f"{name}_lens"]) and np.array_equal(pts, fx[f"{name}_pts"])
        assert np.allclose(np.asarray(an, np.float32), fx[f"{name}_anisotropy"])
