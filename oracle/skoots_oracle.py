"""TEST INFRASTRUCTURE — CPU restatement of the reference's skeleton-embedding
instance-assembly path (buswinka/skoots v0.0.5, `skoots/lib/*.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`skoots_b200/`) never does and has no CPU
fallback.

Parity status: the reference has no tests or golden vectors for this path,
so this oracle is pinned by (1) the reference's two `__main__` known answers
(vector_to_embedding.py:221-232 -> (6,6,6); flood_fill.py:264-277) and (2) outputs of the
*unmodified reference functions run in the build container*, committed as fixtures under
`tests/golden/` by `oracle/gen_golden.py`, and re-checked live by
`tests/test_oracle_vs_reference.py` whenever /root/reference is present.
`scipy.ndimage.label` (scipy 1.18.1 in this image; requirement >=1.10.1,
requirements.txt:36) is the third-party CCL the reference calls (flood_fill.py:135); it is
installed here and on the GPU box and is used as-is.
`bake_skeleton_triton` (the semantics of the reference's Triton kernel) is pinned by outputs of the
unmodified reference's Triton launch on a B200 (`oracle/gen_golden_triton.py` -> tests/golden/bake_triton.npz).
PARITY UNPINNED for two restatements, because the reference code cannot run in this image:
`renumber` (fastremap is absent) and the skeleton-point half of `elastic_deform` (the reference's
`_elastic_on_skeletons` raises under torch 2.11, see oracle/gen_golden_elastic.py).

Each function cites the reference lines it restates. torch CPU ops are used for the bulk
element-wise work so that the timed CPU baseline uses all host threads exactly as the
reference (which is itself torch-on-CPU) would.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Sequence, Tuple

import numpy as np
import torch
from scipy import ndimage as _ndi

# --------------------------------------------------------------------------------------
# a6: crop grid                                                    skoots/lib/cropper.py
# --------------------------------------------------------------------------------------


def clamp_crop(dims: Sequence[int], crop: Sequence[int]) -> List[int]:
    """cropper.py:81-84 — a crop larger than the volume is clamped to it."""
    return [min(int(c), int(d)) for c, d in zip(crop, dims)]


def crop_origins(dim: int, size: int, overlap: int) -> List[int]:
    """Origins visited along one axis (cropper.py:100-142): advance by size-2*overlap while
    the running origin is inside the volume; a crop that would overrun is shifted back to
    dim-size (and may therefore be emitted more than once)."""
    step = size - 2 * overlap
    if step <= 0:
        raise ValueError("crop size must exceed twice the overlap (reference would not terminate)")
    out, o = [], 0
    while o < dim:
        out.append(o if o + size <= dim else dim - size)
        o += step
    return out


def crop_grid(dims, crop, overlap) -> Iterator[Tuple[int, int, int]]:
    """(x,y,z) origins in the reference's loop order: x outermost, z innermost."""
    size = clamp_crop(dims, crop)
    for x in crop_origins(dims[0], size[0], overlap[0]):
        for y in crop_origins(dims[1], size[1], overlap[1]):
            for z in crop_origins(dims[2], size[2], overlap[2]):
                yield x, y, z


# --------------------------------------------------------------------------------------
# a1: vector -> embedding                           skoots/lib/vector_to_embedding.py
# --------------------------------------------------------------------------------------


def vector_to_embedding(scale, vector: torch.Tensor, N: int = 1, decay: float = 1.0) -> torch.Tensor:
    """vector_to_embedding.py:135-174. 5-D -> 3-D walk (79-132); 4-D -> 2-D add (50-76)."""
    s = torch.as_tensor(scale).float()
    if vector.ndim == 4:
        assert N == 1 and decay == 1.0
        B, C, X, Y = vector.shape
        ax = [torch.arange(n, dtype=torch.float32) for n in (X, Y)]
        out = torch.empty((B, 2, X, Y), dtype=torch.float32)
        out[:, 0] = ax[0][:, None] + vector[:, 0].float() * s[0]
        out[:, 1] = ax[1][None, :] + vector[:, 1].float() * s[1]
        return out

    B, C, X, Y, Z = vector.shape
    vf = vector.float()
    ax = (
        torch.arange(X, dtype=torch.float32)[:, None, None],
        torch.arange(Y, dtype=torch.float32)[None, :, None],
        torch.arange(Z, dtype=torch.float32)[None, None, :],
    )
    # :104-105  mesh + vector*scale  (two separately rounded fp32 ops)
    phi = torch.stack([ax[c] + vf[:, c] * s[c] for c in range(3)], dim=1)

    k = 1.0  # python float (double), :107-112
    dims = (X, Y, Z)
    for _ in range(N - 1):
        k *= decay
        ks = torch.tensor(k, dtype=torch.float32) * s  # fp32(k) * s, :114
        idx = [torch.clamp(torch.round(phi[:, c]), 0, dims[c]) for c in range(3)]  # :116-119 (bound is dim, not dim-1)
        flat = (idx[0] * Y * Z) + (idx[1] * Z) + idx[2]  # fp32 ravel, :122-126
        flat = flat.clamp(0, X * Y * Z - 1).long()  # :127
        for c in range(3):
            hop = (vf[:, c] * ks[c]).reshape(-1)[flat.reshape(-1)].reshape(flat.shape)  # take(), :130
            phi[:, c] = phi[:, c] + hop
    return phi


# --------------------------------------------------------------------------------------
# a2: label gather                                      skoots/lib/skeleton.py:656-695
# --------------------------------------------------------------------------------------


def index_skeleton_by_embed(skeleton: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """rint -> clamp to the label volume's [0,dim-1] -> gather -> int32 (skeleton.py:678-695)."""
    assert skeleton.ndim == 5 and embed.ndim == 5
    _, _, x, y, z = embed.shape
    e = torch.round(embed.reshape(3, -1))
    xi = e[0].clamp(0, skeleton.shape[2] - 1).long()
    yi = e[1].clamp(0, skeleton.shape[3] - 1).long()
    zi = e[2].clamp(0, skeleton.shape[4] - 1).long()
    return skeleton[0, 0][xi, yi, zi].to(torch.int32).reshape(1, 1, x, y, z)


# --------------------------------------------------------------------------------------
# a3: connected components of the skeleton mask            skoots/lib/flood_fill.py
# --------------------------------------------------------------------------------------


def label_components(mask: np.ndarray) -> Tuple[np.ndarray, int]:
    """6-connectivity (3-D) / 4-connectivity (2-D) labels numbered by raster order of each
    component's first voxel — scipy.ndimage.label with its default structure
    (flood_fill.py:135; utils/flood_and_stitch.py:63-69)."""
    lab, n = _ndi.label(np.asarray(mask) > 0)
    return lab.astype(np.int32), int(n)


def flood_fill_exact(skeleton: torch.Tensor) -> torch.Tensor:
    """What efficient_flood_fill returns for any volume that fits one 1000x1000x200 crop
    (flood_fill.py:27-50,125-140): scipy labels + 2 on foreground, int16, 3..N+2."""
    vol = skeleton.squeeze(0) if skeleton.ndim == 4 else skeleton
    lab, _ = label_components(vol.numpy())
    out = np.where(lab > 0, lab + 2, 0).astype(np.int16)
    return torch.from_numpy(out)


def _adjacent_by_sum_product(p0: np.ndarray, p1: np.ndarray) -> List[Tuple[int, int]]:
    """flood_fill.py:237-261 — the seam test: labels (a,b) are declared adjacent when a+b and
    a*b (int16 arithmetic) both occur in the element-wise sum / product of the two planes."""
    p0 = p0.astype(np.int16)
    p1 = p1.astype(np.int16)
    with np.errstate(over="ignore"):
        sums = set(np.unique(p0 + p1).tolist())
        prods = set(np.unique(p0 * p1).tolist())
        found = []
        for a in np.unique(p0):
            for b in np.unique(p1):
                if a == 0 or b == 0:
                    continue
                if int(np.int16(a + b)) in sums and int(np.int16(a * b)) in prods:
                    found.append((int(a), int(b)))
    return found


def flood_fill_multicrop(skeleton: torch.Tensor, crop=(1000, 1000, 200)) -> torch.Tensor:
    """Bug-compatible replay of efficient_flood_fill (flood_fill.py:13-122) for volumes larger
    than one crop: per-crop scipy labels offset by the running max (0 after an empty crop,
    :140), seam collisions by the sum/product heuristic, DFS components in insertion order,
    every member replaced by the component's last-visited node.  Not in place."""
    vol = (skeleton.squeeze(0) if skeleton.ndim == 4 else skeleton).clone().to(torch.int16).numpy()
    dims = vol.shape
    size = clamp_crop(dims, crop)
    max_id = 1
    seams: List[List[int]] = [[], [], []]
    for x, y, z in crop_grid(dims, crop, (0, 0, 0)):
        for ax, o in enumerate((x, y, z)):
            if o not in seams[ax]:
                seams[ax].append(o)
        sl = (slice(x, x + size[0]), slice(y, y + size[1]), slice(z, z + size[2]))
        fg = vol[sl] > 0
        lab, _ = _ndi.label(fg)
        with np.errstate(over="ignore"):
            piece = lab.astype(np.int16) + (fg.astype(np.int32) * (max_id + 1)).astype(np.int16)
        vol[sl] = piece
        max_id = int(piece.max())

    pairs: List[Tuple[int, int]] = []
    for ax in range(3):
        for o in seams[ax]:
            if o > 0:
                pairs.extend(_adjacent_by_sum_product(np.take(vol, o, axis=ax), np.take(vol, o - 1, axis=ax)))

    graph: Dict[int, List[int]] = {}
    for a, b in pairs:
        graph.setdefault(a, []).append(b)
        graph.setdefault(b, []).append(a)
    seen = set()
    table: Dict[int, int] = {}
    for start in graph:
        if start in seen:
            continue
        order: List[int] = []
        stack = [(start, iter(graph[start]))]
        seen.add(start)
        order.append(start)
        while stack:  # iterative form of the reference's recursive dfs (flood_fill.py:143-155)
            node, it = stack[-1]
            nxt = next((n for n in it if n not in seen), None)
            if nxt is None:
                stack.pop()
            else:
                seen.add(nxt)
                order.append(nxt)
                stack.append((nxt, iter(graph[nxt])))
        keep = order[-1]
        for member in order[:-1]:
            table.setdefault(member, keep)  # first match wins in the numba loop (:197-203)
    if table:
        lut = np.arange(-32768, 32768, dtype=np.int32)
        for k, v in table.items():
            lut[k + 32768] = v
        vol = lut[vol.astype(np.int32) + 32768].astype(np.int16)
    return torch.from_numpy(vol)


def canonical_relabel(labels) -> np.ndarray:
    """non-zero label -> 1-based rank of its first occurrence in C-order."""
    arr = np.asarray(labels)
    flat = arr.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    keep = uniq != 0
    uniq, first = uniq[keep], first[keep]
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(1, len(uniq) + 1)
    pos = np.searchsorted(uniq, flat)
    pos = np.clip(pos, 0, max(len(uniq) - 1, 0))
    out = np.zeros(flat.shape, dtype=np.int64)
    if len(uniq):
        hit = (flat != 0) & (uniq[pos] == flat)
        out[hit] = rank[pos[hit]]
    return out.reshape(arr.shape)


# --------------------------------------------------------------------------------------
# a4: morphology                                          skoots/lib/morphology.py
# --------------------------------------------------------------------------------------


def _window_reduce(x: torch.Tensor, reach: Tuple[int, int, int], op) -> torch.Tensor:
    """zero-padded window reduction over (B,C,X,Y,Z); the reference extracts the window with a
    one-hot conv3d (zero padding) and reduces over the tap axis (morphology.py:145-152)."""
    rx, ry, rz = reach
    p = torch.nn.functional.pad(x, (rz, rz, ry, ry, rx, rx), value=0.0)
    X, Y, Z = x.shape[-3:]
    acc = None
    for dx in range(2 * rx + 1):
        for dy in range(2 * ry + 1):
            for dz in range(2 * rz + 1):
                tap = p[..., dx:dx + X, dy:dy + Y, dz:dz + Z]
                acc = tap.clone() if acc is None else op(acc, tap)
    return acc


def binary_dilation(image: torch.Tensor) -> torch.Tensor:
    """3x3x3 zero-padded max (morphology.py:155-175)."""
    return _window_reduce(image, (1, 1, 1), torch.maximum)


def binary_dilation_2d(image: torch.Tensor) -> torch.Tensor:
    """3x3x1 zero-padded max (morphology.py:178-199; SURVEY B#11)."""
    return _window_reduce(image, (1, 1, 0), torch.maximum)


def binary_erosion(image: torch.Tensor) -> torch.Tensor:
    """3x3x3 zero-padded min; returns (1, B*C, X, Y, Z) (morphology.py:130-152)."""
    b, c, X, Y, Z = image.shape
    return _window_reduce(image, (1, 1, 1), torch.minimum).reshape(1, b * c, X, Y, Z)


# --------------------------------------------------------------------------------------
# a5: tile epilogue                                        skoots/lib/eval.py:145-176
# --------------------------------------------------------------------------------------


def tile_epilogue(out: torch.Tensor, vectors: torch.Tensor, skeleton: torch.Tensor, origin,
                  overlap=(50, 50, 5), threshold: float = 0.8) -> None:
    """UNet output (1,C>=5,x,y,z) -> masked vectors (fp16) and dilated+thresholded skeleton
    (u8) written into the interior of the tile's slot in the whole-volume arrays."""
    prob = out[:, [-1]]
    skel = out[:, [-2]].float()
    vec = out[:, 0:3]
    keep = prob.gt(threshold)
    vec = vec * keep
    skel = skel * keep
    skel = binary_dilation(skel)
    skel = binary_dilation_2d(binary_dilation_2d(skel))
    x, y, z = origin
    sx, sy, sz = out.shape[2:]
    ox, oy, oz = overlap
    dst = (Ellipsis, slice(x + ox, x + sx - ox), slice(y + oy, y + sy - oy), slice(z + oz, z + sz - oz))
    src = (0, Ellipsis, slice(ox, sx - ox), slice(oy, sy - oy), slice(oz, sz - oz))
    vectors[dst] = vec[src].half()
    skeleton[dst] = skel[src].gt(threshold).to(skeleton.dtype)


# --------------------------------------------------------------------------------------
# a1+a2+a6: instance assembly loop                          skoots/lib/eval.py:245-284
# --------------------------------------------------------------------------------------


def assemble_instances(labels: torch.Tensor, vectors: torch.Tensor, scale, N: int = 10, decay: float = 1.0,
                       crop=(500, 500, 50), overlap=(50, 50, 5), out_dtype=torch.int16) -> torch.Tensor:
    """labels (X,Y,Z) integer label volume, vectors (3,X,Y,Z) -> instance mask (X,Y,Z).
    Per crop: crop-local walk, origin added afterwards in fp32 (:274), gather against the
    whole-volume labels, interior written (later crops overwrite); the outer `overlap`
    margin is never written (:259-269)."""
    dims = tuple(labels.shape)
    size = clamp_crop(dims, crop)
    inst = torch.zeros(dims, dtype=out_dtype)
    lab5 = labels[None, None]
    for x, y, z in crop_grid(dims, crop, overlap):
        v = vectors[:, x:x + size[0], y:y + size[1], z:z + size[2]][None]
        emb = vector_to_embedding(scale, v, N=N, decay=decay)
        emb += torch.tensor((x, y, z), dtype=torch.float32).view(1, 3, 1, 1, 1)
        got = index_skeleton_by_embed(lab5, emb)[0, 0]
        if all(o > 0 for o in overlap):
            ox, oy, oz = overlap
            inst[x + ox:x + size[0] - ox, y + oy:y + size[1] - oy, z + oz:z + size[2] - oz] = \
                got[ox:size[0] - ox, oy:size[1] - oy, oz:size[2] - oz].to(out_dtype)
        else:
            inst[x:x + size[0], y:y + size[1], z:z + size[2]] = got.to(out_dtype)
    return inst


def postprocess(skeleton_mask: torch.Tensor, vectors: torch.Tensor, scale, N: int = 1, decay: float = 1.0,
                crop=None, overlap=(0, 0, 0), out_dtype=torch.int32) -> torch.Tensor:
    """The whole path the headline metric counts: flood fill the u8 skeleton mask, then
    assemble.  ``crop=None`` = the whole volume as one crop (SURVEY §8d, C1 rule)."""
    labels = flood_fill_exact(skeleton_mask.to(torch.int16))
    dims = tuple(labels.shape)
    return assemble_instances(labels, vectors, scale, N=N, decay=decay,
                              crop=dims if crop is None else crop, overlap=overlap, out_dtype=out_dtype)


def postprocess_2d(skeleton_masks: torch.Tensor, vectors: torch.Tensor, scale, label_base: int = 0) -> torch.Tensor:
    """2-D mode (BASELINE configs[4]) slice by slice: `scipy.ndimage.label` of the plane (4-connectivity,
    utils/flood_and_stitch.py:63-69), `_vec2embed2D` (vector_to_embedding.py:50-76), and — the reference has no 2-D
    gather (skeleton.py:671-673 asserts 5-D) — `index_skeleton_by_embed` with Z = 1.  masks (S,X,Y), vectors (S,2,X,Y)."""
    S, X, Y = skeleton_masks.shape
    out = torch.zeros((S, X, Y), dtype=torch.int32)
    for s in range(S):
        lab, _ = label_components(skeleton_masks[s].numpy())
        lab = torch.from_numpy(lab.astype(np.int32))
        lab = torch.where(lab > 0, lab + label_base, lab)
        emb2 = vector_to_embedding(scale, vectors[s][None])                              # (1,2,X,Y)
        emb3 = torch.cat([emb2, torch.zeros((1, 1, X, Y))], dim=1).unsqueeze(-1)         # (1,3,X,Y,1)
        out[s] = index_skeleton_by_embed(lab[None, None, :, :, None], emb3)[0, 0, :, :, 0]
    return out


# --------------------------------------------------------------------------------------
# a7: embedding -> probability                        skoots/lib/embedding_to_prob.py
# --------------------------------------------------------------------------------------


def baked_embed_to_prob(embedding: torch.Tensor, baked: torch.Tensor, sigma: torch.Tensor, eps: float = 1e-16):
    """exp(sum_c (E_c-S_c)^2 / (-2 (sigma_c+eps)^2)) (embedding_to_prob.py:37-49)."""
    sg = -2.0 * (sigma.float() + eps) ** 2
    shape = [1, -1] + [1] * (embedding.ndim - 2)
    q = (embedding - baked) ** 2 / sg.view(shape)
    return torch.exp(q.sum(dim=1, keepdim=True))


# --------------------------------------------------------------------------------------
# a8: bake skeleton (CPU/torch semantics)                skoots/lib/skeleton.py:370-528
# --------------------------------------------------------------------------------------


def average_baked_skeletons(baked: torch.Tensor) -> torch.Tensor:
    """(B,3,X,Y,Z): per channel sum(window)/max(1,count(window>0)), zero padded 3x3x3
    (skeleton.py:18-48; SURVEY B#10)."""
    total = _window_reduce(baked, (1, 1, 1), torch.add)
    count = _window_reduce(baked.gt(0).float(), (1, 1, 1), torch.add)
    count = torch.where(count == 0, torch.ones_like(count), count)
    return total / count


def bake_skeleton(masks: torch.Tensor, skeletons: Dict[int, torch.Tensor], anisotropy=(1.0, 1.0, 1.0),
                  average: bool = True) -> torch.Tensor:
    """For every voxel of object k: the point of skeleton k nearest to it, anisotropy scaling
    the coordinates, first minimum in skeleton order on ties (skeleton.py:416-443), then the
    masked 27-mean (:519).  Distances are evaluated directly (no matmul expansion): identical
    to the reference's cdist for integer-valued coordinates, where every term is exact."""
    if -1 in skeletons:
        return torch.zeros((3,) + tuple(masks.shape[-3:]), dtype=torch.float16)
    vol = masks.squeeze(0) if masks.ndim == 4 else masks
    X, Y, Z = vol.shape
    baked = torch.zeros((3, X, Y, Z), dtype=torch.float32)
    an = torch.tensor(anisotropy, dtype=torch.float32)
    for k in torch.unique(vol).tolist():
        if k == 0:
            continue
        where = (vol == k).nonzero()
        pts = skeletons[int(k)].float()
        d = ((pts[:, None, :] * an - where[None, :, :].float() * an) ** 2).sum(-1).clamp_min(0).sqrt()
        pick = d.argmin(dim=0)
        baked[:, where[:, 0], where[:, 1], where[:, 2]] = pts[pick].T
    if average:
        baked = average_baked_skeletons(baked[None])[0]
    return baked


def bake_skeleton_triton(masks: torch.Tensor, skeletons: Dict[int, torch.Tensor], anisotropy=(1.0, 1.0, 1.0),
                         average: bool = True):
    """What the reference's TRITON kernel returns for a CUDA mask (skeleton.py:51-367 via :505-512), restated on the CPU:
    dist = sum_c (s_c - v_c)^2 * a_c (:208-212); the skeleton row is loaded SKEL_BLOCK_SIZE wide (:204-206, block =
    next power of two of the longest skeleton, :361) and the lanes past its length hold zeros, i.e. a phantom point at
    the origin; each coordinate = max over the tied lanes and over 0 from the untied ones (:219-221); an id without a
    skeleton uses index 0 with length 0 (:176,:187) -> zeros; fp16 stores.  Returns (baked, distance): baked (3,X,Y,Z)
    fp16, or fp32 after the averaging (:519-523); distance (1,X,Y,Z) fp16 (exact sqrt here, tl.sqrt on the device: the
    fp16 value may differ by one ulp).  PINNED by tests/golden/bake_triton.npz (reference outputs from a B200) for
    integer-valued coordinates and anisotropy."""
    vol = (masks.squeeze(0) if masks.ndim == 4 else masks).numpy()
    X, Y, Z = vol.shape
    baked = np.zeros((3, X, Y, Z), dtype=np.float32)
    dist = np.zeros((1, X, Y, Z), dtype=np.float32)
    longest = max((int(v.shape[0]) for v in skeletons.values()), default=0)
    if longest == 0:
        return torch.zeros((3, X, Y, Z), dtype=torch.float16), torch.zeros((3, X, Y, Z), dtype=torch.float16)
    block = 1 if longest == 0 else 2 ** (longest - 1).bit_length()
    an = np.asarray(anisotropy, dtype=np.float32)
    for k in np.unique(vol).tolist():
        if k == 0:
            continue
        where = np.argwhere(vol == k)
        pts = skeletons[int(k)].float().numpy().reshape(-1, 3) if int(k) in skeletons else np.zeros((0, 3), np.float32)
        lanes = np.zeros((block, 3), dtype=np.float32)
        lanes[:len(pts)] = pts
        e = lanes[:, None, :] - where[None, :, :].astype(np.float32)
        d2 = ((e[..., 0] * e[..., 0]) * an[0] + (e[..., 1] * e[..., 1]) * an[1]) + (e[..., 2] * e[..., 2]) * an[2]
        tied = d2 == d2.min(axis=0, keepdims=True)
        close = np.where(tied[..., None], lanes[:, None, :], np.float32(0)).max(axis=0)
        baked[:, where[:, 0], where[:, 1], where[:, 2]] = close.T
        dist[0, where[:, 0], where[:, 1], where[:, 2]] = np.sqrt(d2.min(axis=0))
    baked_t = torch.from_numpy(baked).to(torch.float16)
    dist_t = torch.from_numpy(dist).to(torch.float16)
    if average:
        return average_baked_skeletons(baked_t[None].float())[0], dist_t
    return baked_t, dist_t


# --------------------------------------------------------------------------------------
# a9: skeleton -> mask                 skoots/lib/skeleton.py:531-593, utils.py:421-438
# --------------------------------------------------------------------------------------


def disk_stamp_offsets(radius: int = 7, flank_radius: int = 3) -> np.ndarray:
    """(S,3) integer offsets: disk(radius) at dz=0, disk(flank) at dz=+-1, x/y shifted by
    -radius//2 (sic, utils.py:435-436), in torch.nonzero (row-major) order."""
    def disk(r):
        span = np.arange(-r, r + 1)
        xx, yy = np.meshgrid(span, span)
        return (xx * xx + yy * yy) <= r * r
    centre, flank = disk(radius), disk(flank_radius)
    pad = (centre.shape[0] - flank.shape[0]) // 2
    flank = np.pad(flank, pad)
    stack = np.stack((flank, centre, flank), axis=-1)
    off = np.argwhere(stack).astype(np.int64)
    off[:, 2] -= 1
    off[:, :2] -= radius // 2
    return off


def skeleton_to_mask(skeletons: Dict[int, torch.Tensor], shape, radius: int = 7, flank_radius: int = 3):
    """OR-stamp around every skeleton point; float coords are added to the offsets and then
    truncated toward zero (`.long()`, skeleton.py:563-569)."""
    if -1 in skeletons:
        return torch.zeros(tuple(shape))
    out = torch.zeros(tuple(shape), dtype=torch.float32)
    off = torch.from_numpy(disk_stamp_offsets(radius, flank_radius))
    for pts in skeletons.values():
        pos = (pts.T.unsqueeze(1) + off.T.unsqueeze(2)).reshape(3, -1).long()
        ok = (pos[0] >= 0) & (pos[0] < shape[0]) & (pos[1] >= 0) & (pos[1] < shape[1]) & \
             (pos[2] >= 0) & (pos[2] < shape[2])
        out[pos[0, ok], pos[1, ok], pos[2, ok]] = 1.0
    return out.unsqueeze(0)


# --------------------------------------------------------------------------------------
# f2: renumber + validation metrics      skoots/lib/eval.py:304, skoots/validate/lib.py
# --------------------------------------------------------------------------------------
def renumber(labels) -> Tuple[np.ndarray, Dict[int, int]]:
    """eval.py:304 `fastremap.renumber(mask, in_place=True)`.  fastremap (seung-lab/fastremap, unpinned in the
    reference's requirements and absent from this image) is third-party: PARITY UNPINNED for it.  Its published
    behaviour is restated: labels are replaced by 1..N in order of first appearance in memory (C) order,
    0 is preserved, and the old->new mapping is returned."""
    arr = np.asarray(labels)
    out = canonical_relabel(arr)
    flat_old, flat_new = arr.reshape(-1), out.reshape(-1)
    _, first = np.unique(flat_old, return_index=True)
    remap = {int(flat_old[i]): int(flat_new[i]) for i in first}
    remap.setdefault(0, 0)
    return out.astype(arr.dtype), remap


def _contingency(gt, pred):
    g, p = np.asarray(gt).reshape(-1).astype(np.int64), np.asarray(pred).reshape(-1).astype(np.int64)
    a_unique = np.unique(g)
    a_unique = a_unique[a_unique > 0]                     # validate/lib.py:201-202
    b_unique = np.unique(p)
    b_unique = b_unique[b_unique > 0]                     # validate/lib.py:204-205
    ia = np.searchsorted(a_unique, g)
    ib = np.searchsorted(b_unique, p)
    ok_a = (g > 0)
    ok_b = (p > 0)
    area_a = np.bincount(ia[ok_a], minlength=len(a_unique)).astype(np.int64)
    area_b = np.bincount(ib[ok_b], minlength=len(b_unique)).astype(np.int64)
    both = ok_a & ok_b
    inter = np.zeros((len(a_unique), len(b_unique)), dtype=np.int64)
    if both.any():
        np.add.at(inter, (ia[both], ib[both]), 1)
    return a_unique, b_unique, inter, area_a, area_b


def mask_iou(gt, pred) -> torch.Tensor:
    """validate/lib.py:190-229.  The reference loops over object pairs; every entry is
    `logical_and(_a,_b).sum() / logical_or(_a,_b).sum()` — a division of two 0-dim int64 tensors, i.e. both are
    converted to float32 and divided once — or 0.0 when the two objects do not touch (:214-227)."""
    _, _, inter, area_a, area_b = _contingency(gt, pred)
    union = area_a[:, None] + area_b[None, :] - inter
    out = torch.zeros(inter.shape, dtype=torch.float32)
    hit = inter > 0
    if hit.any():
        num = torch.from_numpy(inter[hit]).to(torch.float32)
        den = torch.from_numpy(union[hit]).to(torch.float32)
        out[torch.from_numpy(hit)] = num / den
    return out


def mask_dice(gt, pred) -> torch.Tensor:
    """validate/lib.py:232-275: 2·|a∩b| / (|a|+|b|) with the same int64 -> float32 division; raises
    AssertionError like the reference (:266-268) when numerator >= denominator (two identical objects)."""
    _, _, inter, area_a, area_b = _contingency(gt, pred)
    den = area_a[:, None] + area_b[None, :]
    hit = inter > 0
    assert not (hit & (2 * inter >= den)).any(), "numerator >= denominator"
    out = torch.zeros(inter.shape, dtype=torch.float32)
    if hit.any():
        out[torch.from_numpy(hit)] = torch.from_numpy((2 * inter)[hit]).to(torch.float32) / torch.from_numpy(
            np.broadcast_to(den, inter.shape)[hit].copy()).to(torch.float32)
    return out


def accuracies_from_iou(iou: torch.Tensor, thr: float = 0.1) -> Tuple[int, int, int]:
    """validate/lib.py:170-187."""
    gt_hit = iou.max(dim=1)[0].gt(thr)
    pred_hit = iou.max(dim=0)[0].gt(thr)
    return int(gt_hit.sum()), int((~pred_hit).sum()), int((~gt_hit).sum())



# --------------------------------------------------------------------------------------
# f4: elastic deformation of the augmentation     skoots/train/merged_transform.py:75-188, 43-72
# --------------------------------------------------------------------------------------


def elastic_deform(noise: torch.Tensor, *args: torch.Tensor, skeleton: Dict[int, torch.Tensor],
                   displacement_magnitude=(0.05, 0.05, 0.01)):
    """The reference's operations with the random field injected: `noise` is what its
    `torch.rand((1, 3, ds[2], ds[1], ds[0]))` (:140) returns.  Trilinear upsampling to the crop, identity grid from
    three linspaces, nearest grid_sample of every argument, then the second grid ((base - offset + 1)/2 * size) looked
    up at every in-volume skeleton point, components reversed, assigned into the (integer) point tensor."""
    import torch.nn.functional as F
    b, c, x, y, z = args[0].shape
    mag = torch.tensor(tuple(reversed(displacement_magnitude)), dtype=torch.float32)
    offset = F.interpolate(noise.float(), (x, y, z), mode="trilinear").permute((0, 2, 3, 4, 1)).mul(mag.view(1, 1, 1, 1, 3))
    d1, d2, d3 = torch.linspace(-1, 1, x), torch.linspace(-1, 1, y), torch.linspace(-1, 1, z)
    meshx, meshy, meshz = torch.meshgrid((d1, d2, d3), indexing="ij")
    base = torch.stack((meshz, meshy, meshx), 3).unsqueeze(0)
    grid = (base + offset).float()
    out = [F.grid_sample(a.float(), grid, align_corners=True, mode="nearest") for a in args]
    grid = (base - offset).float().add(1).div(2).mul(torch.tensor((z, y, x)).view(1, 1, 1, 1, 3))
    new = {}
    for k, skel in skeleton.items():
        skel = skel.clone()
        sx, sy, sz = skel[:, 0], skel[:, 1], skel[:, 2]
        ind = (sx >= 0) & (sx < x) & (sy >= 0) & (sy < y) & (sz >= 0) & (sz < z)
        skel[ind, :] = grid[0, sx[ind].long(), sy[ind].long(), sz[ind].long(), :][:, [2, 1, 0]].to(skel.dtype)
        new[k] = skel
    return (*out, new)
