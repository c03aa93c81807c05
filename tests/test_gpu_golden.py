"""GPU: the CUDA path (through the ctypes C-ABI) against the fixtures produced by the unmodified
reference.  Integer outputs bit-exact; embeddings bit-exact (the kernel restates every fp32
rounding of the reference)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, unpack_mask

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def test_library_loaded_is_in_tree():
    import skoots_b200._lib as L
    assert L.load().skb_version() == 100
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
