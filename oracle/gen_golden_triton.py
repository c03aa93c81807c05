"""TEST INFRASTRUCTURE — tests/golden/bake_triton.npz from the UNMODIFIED reference's Triton `bake_skeleton`
(skoots/lib/skeleton.py:51-367, dispatched for CUDA masks at :505-512).  It needs a GPU, so unlike the other generators
it runs on a B200 box, against the copy of the reference's python sources that `oracle/build_ref.py` keeps under
`oracle/_ref` (git-ignored; it travels with gpurun):

    gpurun -- 'python oracle/gen_golden_triton.py gpurun_out/bake_triton.npz'      # then: mv into tests/golden/

The inputs are the seeded cases of `cases()` below; the outputs are what the reference returned on that box:
raw fp16 baked points and fp16 distances (average=False, return_distance=True) and the averaged fp32 field
(average=True).  The script also asserts that the oracle restatement (skoots_oracle.bake_skeleton_triton) equals them,
and, when the CUDA library is there, prints how the product's triton_compat mode compares.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch


def _blobs(shape, centres, radius, ids):
    """voxel -> id of the nearest centre within `radius` (0 elsewhere)."""
    grid = np.stack(np.meshgrid(*[np.arange(n) for n in shape], indexing="ij"), -1).astype(np.float32)
    d = np.linalg.norm(grid[..., None, :] - np.asarray(centres, np.float32), axis=-1)
    near = d.argmin(-1)
    return np.where(d.min(-1) <= radius, np.asarray(ids)[near], 0).astype(np.int32)


def cases():
    """name -> (mask int32 (X,Y,Z), {id: int64 (M,3)}, anisotropy).  Integer coordinates and anisotropy: the pinned domain."""
    out = {}
    rng = np.random.default_rng(20261018)

    # several objects, skeleton lengths 1..16 (the 16-point one fills the block: no phantom lanes for it), symmetric
    # pairs of points (ties), non-sequential ids
    shape = (40, 36, 20)
    ids = [3, 234, 6, 17, 9, 1000]
    centres = [(8, 8, 6), (30, 9, 10), (10, 27, 12), (30, 28, 5), (20, 18, 15), (4, 3, 2)]
    lens = [1, 5, 8, 13, 16, 3]
    sk = {}
    for k, c, n in zip(ids, centres, lens):
        pts = np.asarray(c)[None] + rng.integers(-4, 5, size=(n, 3))
        if n >= 5:  # mirror images about the centre: equidistant from voxels on the mid-plane
            pts[1] = np.asarray(c) + (2, 0, 0)
            pts[2] = np.asarray(c) - (2, 0, 0)
            pts[3] = np.asarray(c) + (0, 2, 1)
            pts[4] = np.asarray(c) - (0, 2, 1)
        sk[k] = torch.from_numpy(np.clip(pts, 0, np.asarray(shape) - 1).astype(np.int64))
    out["blobs"] = (_blobs(shape, centres, 7.5, ids), sk, (1.0, 1.0, 1.0))

    # anisotropy, and an object hugging the origin whose skeleton is far away: the phantom origin point wins there
    shape = (32, 32, 16)
    ids = [5, 12]
    centres = [(3, 3, 2), (20, 20, 8)]
    sk = {5: torch.tensor([[14, 3, 2], [3, 14, 2], [9, 9, 6]]), 12: torch.tensor([[20, 20, 8], [22, 20, 8], [18, 20, 8], [20, 24, 9], [20, 16, 7]])}
    out["aniso"] = (_blobs(shape, centres, 6.0, ids), sk, (1.0, 1.0, 5.0))

    # an id in the mask without a skeleton (zeros, no error), next to a normal object
    shape = (24, 24, 12)
    mask = _blobs(shape, [(6, 6, 5), (17, 16, 6)], 5.0, [77, 8])
    out["missing"] = (mask, {8: torch.tensor([[17, 16, 6], [15, 16, 6], [19, 16, 6], [17, 18, 7]])}, (1.0, 1.0, 3.0))

    # a training-crop-like case: 12 objects, up to 60 points (block 64)
    shape = (96, 96, 20)
    ids = list(range(1, 13))
    centres = [tuple(int(v) for v in rng.integers((8, 8, 3), (88, 88, 17))) for _ in ids]
    sk = {}
    for k, c in zip(ids, centres):
        n = int(rng.integers(4, 61))
        walk = np.cumsum(rng.integers(-1, 2, size=(n, 3)), 0) + np.asarray(c)
        sk[k] = torch.from_numpy(np.clip(walk, 0, np.asarray(shape) - 1).astype(np.int64))
    out["crop"] = (_blobs(shape, centres, 9.0, ids), sk, (1.0, 1.0, 5.0))
    return out


def pack_skeletons(sk):
    keys = list(sk.keys())
    return (np.asarray(keys, np.int64), np.asarray([int(sk[k].shape[0]) for k in keys], np.int64),
            np.concatenate([sk[k].numpy().reshape(-1, 3) for k in keys], 0).astype(np.int64))


def unpack_skeletons(ids, lens, pts):
    out, at = {}, 0
    for k, n in zip(ids.tolist(), lens.tolist()):
        out[int(k)] = torch.from_numpy(pts[at:at + n].copy())
        at += n
    return out


def main():
    dest = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "bake_triton.npz")
    import ref_shim
    ref_shim.install(need_morphology=False)
    import skoots.lib.skeleton as ref
    import skoots_oracle as orc
    dev = torch.device("cuda:0")
    pack, report = {}, []
    for name, (mask, sk, an) in cases().items():
        m = torch.from_numpy(mask).to(dev)
        skd = {k: v.to(dev).float().contiguous() for k, v in sk.items()}
        raw, dist = ref.bake_skeleton(m, skd, an, average=False, return_distance=True)
        avg = ref.bake_skeleton(m, skd, an, average=True)
        assert raw.dtype == torch.float16 and dist.dtype == torch.float16 and avg.dtype == torch.float32
        ids, lens, pts = pack_skeletons(sk)
        pack.update({f"{name}_mask": mask, f"{name}_ids": ids, f"{name}_lens": lens, f"{name}_pts": pts,
                     f"{name}_anisotropy": np.asarray(an, np.float32), f"{name}_raw": raw.cpu().numpy(),
                     f"{name}_dist": dist.cpu().numpy(), f"{name}_avg": avg.cpu().numpy()})
        o_raw, o_dist = orc.bake_skeleton_triton(torch.from_numpy(mask), sk, an, average=False)
        o_avg, _ = orc.bake_skeleton_triton(torch.from_numpy(mask), sk, an, average=True)
        ulp = (o_dist.view(torch.int16).int() - dist.cpu().view(torch.int16).int()).abs().max().item()
        line = {"case": name, "voxels": int(mask.size), "foreground": int((mask != 0).sum()),
                "oracle_raw_equal": bool(torch.equal(o_raw, raw.cpu())), "oracle_dist_max_ulp": int(ulp),
                "oracle_avg_max_abs": float((o_avg - avg.cpu()).abs().max())}
        try:
            from skoots_b200.lib import skeleton as mine
            g_raw, g_dist = mine.bake_skeleton(m, skd, an, average=False, return_distance=True, triton_compat=True)
            g_avg = mine.bake_skeleton(m, skd, an, average=True, triton_compat=True)
            line.update({"b200_raw_equal": bool(torch.equal(g_raw, raw)), "b200_raw_mismatches": int((g_raw != raw).any(0).sum()),
                         "b200_dist_max_ulp": int((g_dist.view(torch.int16).int() - dist.view(torch.int16).int()).abs().max()),
                         "b200_avg_max_abs": float((g_avg - avg).abs().max())})
        except Exception as exc:  # the fixture does not depend on the product
            line["b200"] = repr(exc)[:200]
        report.append(line)
        print(line, flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(dest)), exist_ok=True)
    np.savez_compressed(dest, **pack)
    print(f"bake_triton: {os.path.getsize(dest) / 1024:.1f} KiB -> {dest}")
    assert all(r["oracle_raw_equal"] and r["oracle_dist_max_ulp"] <= 1 and r["oracle_avg_max_abs"] <= 1e-4 for r in report), \
        "oracle restatement differs from the reference's Triton kernel"


if __name__ == "__main__":
    main()
