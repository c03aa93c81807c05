"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
functions (imported from /root/reference through oracle/ref_shim.py) on small seeded inputs.

    PYTORCH_JIT=0 python oracle/gen_golden.py

The reference cannot travel to the GPU box, so these fixtures are what pins both the oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py) to the reference's
own outputs.  `skoots.lib.eval.eval` itself cannot run here (zarr/fastremap/bism/checkpoint
are absent): its post-UNet sections are replayed below *around the reference's functions*,
line for line in meaning (eval.py:145-176 and :245-284), with the crop sizes as parameters.
"""
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import contextlib
import io

import numpy as np
import torch

import ref_shim

ref_shim.install()

from skoots.lib.cropper import crops as ref_crops  # noqa: E402
from skoots.lib.embedding_to_prob import baked_embed_to_prob as ref_prob  # noqa: E402
from skoots.lib.flood_fill import connected_components as ref_graph_cc  # noqa: E402
from skoots.lib.flood_fill import efficient_flood_fill as ref_flood  # noqa: E402
from skoots.lib.morphology import binary_dilation as ref_dil  # noqa: E402
from skoots.lib.morphology import binary_dilation_2d as ref_dil2d  # noqa: E402
from skoots.lib.morphology import binary_erosion as ref_ero  # noqa: E402
from skoots.lib.skeleton import average_baked_skeletons as ref_avg  # noqa: E402
from skoots.lib.skeleton import bake_skeleton as ref_bake  # noqa: E402
from skoots.lib.skeleton import index_skeleton_by_embed as ref_index  # noqa: E402
from skoots.lib.skeleton import skeleton_to_mask as ref_s2m  # noqa: E402
from skoots.lib.utils import get_cached_disk_coords as ref_disk  # noqa: E402
from skoots.lib.vector_to_embedding import vector_to_embedding as ref_v2e  # noqa: E402

from skoots_b200.synthetic import make_tube_volume  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def ref_assemble(labels_i16, vectors, scale, N, decay, crop, overlap):
    """eval.py:245-284 around the reference functions (crop/overlap parametrised)."""
    inst = torch.zeros_like(labels_i16, dtype=torch.int16)
    skeleton = labels_i16.unsqueeze(0).unsqueeze(0)
    cropsize = list(crop)
    for _vec, (x, y, z) in ref_crops(vectors, crop_size=cropsize, overlap=overlap):
        dst = (slice(x + overlap[0], x + cropsize[0] - overlap[0]),
               slice(y + overlap[1], y + cropsize[1] - overlap[1]),
               slice(z + overlap[2], z + cropsize[2] - overlap[2]))
        src = (slice(overlap[0], -overlap[0]), slice(overlap[1], -overlap[1]), slice(overlap[2], -overlap[2]))
        emb = ref_v2e(scale=scale, vector=_vec, N=N, decay=decay)
        emb += torch.tensor((x, y, z)).view(1, 3, 1, 1, 1)
        got = ref_index(skeleton=skeleton, embed=emb).squeeze()
        inst[dst] = got[src] if torch.tensor(overlap).gt(0).all() else got
    return inst


def ref_tile_epilogue(out, vectors, skeleton, origin, overlap, cropsize):
    """eval.py:145-176 around the reference functions."""
    x, y, z = origin
    prob = out[:, [-1], ...]
    skel = out[:, [-2], ...].float()
    vec = out[:, 0:3:1, ...]
    vec = vec * prob.gt(0.8)
    skel = skel * prob.gt(0.8)
    skel = ref_dil(skel)
    for _ in range(2):
        skel = ref_dil2d(skel)
    dst = (..., slice(x + overlap[0], x + cropsize[0] - overlap[0]),
           slice(y + overlap[1], y + cropsize[1] - overlap[1]),
           slice(z + overlap[2], z + cropsize[2] - overlap[2]))
    src = (0, ..., slice(overlap[0], -overlap[0]), slice(overlap[1], -overlap[1]), slice(overlap[2], -overlap[2]))
    vectors[dst] = vec[src].half()
    skeleton[dst] = skel[src].gt(0.8).to(torch.uint8)


def gen_2d(g=None):
    """a10, 2-D mode (BASELINE configs[4]): per-slice scipy label (utils/flood_and_stitch.py:63-69), the reference's
    2-D vector_to_embedding and — it has no 2-D gather — its index_skeleton_by_embed applied per slice with Z = 1."""
    from scipy.ndimage import label as ndi_label
    g = g or torch.Generator().manual_seed(4321)
    tv = make_tube_volume((3, 72, 64), 30, seed=9, flat=True, scale=(1.0, 12.0, 12.0))
    noise = (torch.rand((3, 72, 64), generator=g) < 0.03).to(torch.uint8)
    masks = torch.maximum(tv.skeleton, noise)
    vec = tv.vectors[1:3].permute(1, 0, 2, 3).contiguous().float()     # (S,2,X,Y): the in-plane components
    vec[0, :, :8] *= 6.0                                                 # targets far outside the image: clamp to the border
    scale = torch.tensor((12.0, 12.0))
    out = np.zeros((3, 72, 64), dtype=np.int32)
    for s_ in range(3):
        plane = (masks[s_].numpy() > 0).astype(np.int32)
        ndi_label(plane, output=plane)                                   # flood_and_stitch.py:66-69
        emb2 = ref_v2e(scale, vec[s_][None])
        emb3 = torch.cat([emb2, torch.zeros((1, 1, 72, 64))], dim=1).unsqueeze(-1)
        lab5 = torch.from_numpy(plane)[None, None, :, :, None]
        out[s_] = ref_index(lab5, emb3)[0, 0, :, :, 0].numpy()
    save("assembly_2d", masks=np.packbits(masks.numpy()), shape=np.array(masks.shape), vectors=vec.numpy(), scale=scale.numpy(), out=out)


def main():
    g = torch.Generator().manual_seed(1234)

    # ---- known answers the reference states itself -------------------------------------
    v = torch.ones((1, 3, 10, 10, 10))
    v[:, :, 5, 5, 5] = -1
    v[:, [0, 1, 2], 4, 4, 4] = torch.tensor((2.0, 2.0, 2.0))
    kat = ref_v2e(torch.tensor((1, 1, 1)), v, N=2)
    assert kat[0, :, 5, 5, 5].tolist() == [6.0, 6.0, 6.0]  # vector_to_embedding.py:221-232
    graph = {1: [2, 3], 2: [1], 3: [1, 5, 4], 4: [5], 5: [3], 6: [7], 7: [6, 8, 9], 8: [7], 9: [7]}
    assert ref_graph_cc(graph) == [[1, 2, 3, 5, 4], [6, 7, 8, 9]]  # flood_fill.py:264-277
    save("kat_vec2embed", vector=v.numpy(), scale=np.array([1, 1, 1]), N=2, out=kat.numpy())

    # ---- a1: vector_to_embedding, random out-of-range fields -----------------------------
    for tag, dt in (("f16", torch.float16), ("bf16", torch.bfloat16), ("f32", torch.float32)):
        X, Y, Z = 22, 18, 12
        vec = (torch.rand((1, 3, X, Y, Z), generator=g) * 2 - 1)
        vec[0, :, :4] *= 3.0  # |v*s| far outside the volume: exercises clamp-to-dim and ravel bleed
        vec = vec.to(dt)
        scale = torch.tensor((9, 7, 5))
        pack = dict(vector=vec.float().numpy(), scale=scale.numpy())
        for N, decay in ((1, 1.0), (2, 1.0), (5, 1.0), (10, 1.0), (10, 0.95)):
            pack[f"out_N{N}_d{int(decay * 100)}"] = ref_v2e(scale, vec, N=N, decay=decay).numpy()
        save(f"vec2embed_{tag}", **pack)
    vec2 = (torch.rand((2, 2, 33, 17), generator=g) * 2 - 1)
    save("vec2embed_2d", vector=vec2.numpy(), scale=np.array([11.0, 4.0], dtype=np.float32),
         out=ref_v2e(torch.tensor((11.0, 4.0)), vec2).numpy())

    # ---- a2: index_skeleton_by_embed ----------------------------------------------------------
    lab = torch.randint(0, 300, (1, 1, 17, 13, 9), generator=g).to(torch.int16)
    emb = torch.rand((1, 3, 11, 10, 7), generator=g) * torch.tensor((24.0, 20.0, 14.0)).view(1, 3, 1, 1, 1) - 3.0
    emb[0, :, 0, 0, :] = torch.tensor([0.5, 1.5, 2.5, -0.5, 3.5, 4.5, 8.5])  # half-to-even cases
    save("index_by_embed", labels=lab.numpy(), embed=emb.numpy(), out=ref_index(lab, emb).numpy())

    # ---- a3: flood fill ---------------------------------------------------------------------------
    tv = make_tube_volume((64, 48, 20), 14, seed=3)
    noise = (torch.rand((64, 48, 20), generator=g) < 0.04).to(torch.uint8)
    m = torch.maximum(tv.skeleton, noise)
    out = quiet(ref_flood, m.to(torch.int16).clone())
    save("flood_small", mask=np.packbits(m.numpy()), shape=np.array(m.shape), out=out.numpy())
    dense = (torch.rand((40, 36, 28), generator=g) < 0.45).to(torch.uint8)  # many merges, worst case for union-find
    out = quiet(ref_flood, dense.to(torch.int16).clone())
    save("flood_dense", mask=np.packbits(dense.numpy()), shape=np.array(dense.shape), out=out.numpy())
    # multi-crop (two 1000-wide crops along x, the second shifted to 100): tubes crossing the seam
    tvm = make_tube_volume((1100, 40, 24), 60, seed=5)
    out = quiet(ref_flood, tvm.skeleton.to(torch.int16).clone())
    save("flood_multicrop", mask=np.packbits(tvm.skeleton.numpy()), shape=np.array(tvm.skeleton.shape),
         out=out.numpy())

    # ---- a4: morphology -------------------------------------------------------------------------------
    img = torch.randn((2, 2, 13, 11, 9), generator=g)
    save("morphology", image=img.numpy(), dilation=ref_dil(img).numpy(), dilation_2d=ref_dil2d(img).numpy(),
         erosion=ref_ero(img).numpy())
    imgb = (torch.rand((1, 1, 20, 20, 6), generator=g) > 0.7).float()
    save("morphology_binary", image=imgb.numpy(), dilation=ref_dil(imgb).numpy(),
         dilation_2d=ref_dil2d(imgb).numpy(), erosion=ref_ero(imgb).numpy())

    # ---- a5: tile epilogue -----------------------------------------------------------------------------
    unet = torch.rand((1, 5, 30, 28, 12), generator=g)
    unet[:, 0:3] = unet[:, 0:3] * 2 - 1
    vol_v = torch.zeros((3, 40, 40, 16), dtype=torch.float16)
    vol_s = torch.zeros((1, 40, 40, 16), dtype=torch.uint8)
    ref_tile_epilogue(unet, vol_v, vol_s, (6, 9, 3), (5, 4, 2), (30, 28, 12))
    save("tile_epilogue", unet=unet.numpy(), origin=np.array((6, 9, 3)), overlap=np.array((5, 4, 2)),
         vectors=vol_v.float().numpy(), skeleton=vol_s.numpy())

    # ---- a1+a2+a6: assembly ----------------------------------------------------------------------------
    tva = make_tube_volume((70, 60, 24), 12, seed=7, scale=(9.0, 9.0, 4.0))
    labels = quiet(ref_flood, tva.skeleton.to(torch.int16).clone())
    scale = torch.tensor((9, 9, 4))
    pack = dict(skeleton=np.packbits(tva.skeleton.numpy()), shape=np.array(tva.skeleton.shape),
                vectors=tva.vectors.float().numpy(), scale=scale.numpy(), labels=labels.numpy())
    for name, N, decay, crop, ov in (("N10", 10, 1.0, (40, 40, 16), (5, 5, 2)),
                                     ("N10d95", 10, 0.95, (40, 40, 16), (5, 5, 2)),
                                     ("N1", 1, 1.0, (40, 40, 16), (5, 5, 2)),
                                     ("N3big", 3, 1.0, (500, 500, 50), (6, 6, 3))):
        pack["inst_" + name] = ref_assemble(labels, tva.vectors, scale, N, decay, crop, ov).numpy()
        pack["cfg_" + name] = np.array([N, int(decay * 100), *crop, *ov])
    # whole-volume N=1 rule for C1: lib functions directly, no margins
    emb = ref_v2e(scale, tva.vectors[None], N=1)
    pack["inst_whole_N1"] = ref_index(labels[None, None], emb)[0, 0].numpy()
    emb = ref_v2e(scale, tva.vectors[None], N=4)
    pack["inst_whole_N4"] = ref_index(labels[None, None], emb)[0, 0].numpy()
    save("assembly", **pack)

    gen_2d(g)

    # ---- a7: baked_embed_to_prob ------------------------------------------------------------------------
    E = torch.rand((2, 3, 9, 8, 7), generator=g) * 30
    S = torch.rand((2, 3, 9, 8, 7), generator=g) * 30
    sig = torch.tensor((20.0, 20.0, 6.0))
    save("embed_prob", embedding=E.numpy(), baked=S.numpy(), sigma=sig.numpy(), out=ref_prob(E, S, sig).numpy())

    # ---- a8: bake_skeleton (CPU semantics) ------------------------------------------------------------------
    tvb = make_tube_volume((48, 40, 12), 8, seed=11)
    present = {int(k): tvb.skeletons[int(k)] for k in torch.unique(tvb.mask).tolist() if k != 0}
    ids = np.array(sorted(present))
    pts = np.concatenate([present[int(k)].numpy() for k in ids])
    lens = np.array([present[int(k)].shape[0] for k in ids])
    pack = dict(mask=tvb.mask.numpy(), ids=ids, lens=lens, points=pts)
    for tag, an in (("iso", (1.0, 1.0, 1.0)), ("aniso", (1.0, 1.0, 3.0))):
        pack[f"baked_{tag}"] = ref_bake(tvb.mask, present, anisotropy=an, average=False).numpy()
        pack[f"baked_avg_{tag}"] = ref_bake(tvb.mask, present, anisotropy=an, average=True).numpy()
    save("bake_skeleton", **pack)
    b = torch.rand((1, 3, 9, 8, 7), generator=g) * (torch.rand((1, 3, 9, 8, 7), generator=g) > 0.5)
    save("average_baked", baked=b.numpy(), out=ref_avg(b).numpy())

    # ---- a9: skeleton_to_mask -----------------------------------------------------------------------------
    sk = {1: torch.tensor([[5.0, 6.0, 3.0], [20.7, 3.2, 0.9], [-2.5, 10.0, 5.0]]),
          2: torch.tensor([[30.0, 30.0, 7.0], [39.0, 1.0, 0.0]])}
    pack = dict(points=np.concatenate([v.numpy() for v in sk.values()]), lens=np.array([3, 2]))
    for r, f in ((7, 3), (9, 3), (2, 1)):
        pack[f"mask_r{r}_f{f}"] = ref_s2m(sk, (40, 36, 8), radius=r, flank_radius=f).numpy()
        pack[f"offsets_r{r}_f{f}"] = ref_disk("cpu", r, f).numpy()
    save("skeleton_to_mask", **pack)


if __name__ == "__main__":
    main()
