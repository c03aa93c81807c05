"""TEST INFRASTRUCTURE — the reference's OWN post-processing path, run unmodified on host cores.

bench.py's `cpu_baseline` leg and `--impl reference` arm call `postprocess` below; tests use it as the
live checker.  The functions executed are the reference's (`skoots.lib.flood_fill.efficient_flood_fill`,
`skoots.lib.vector_to_embedding.vector_to_embedding`, `skoots.lib.skeleton.index_skeleton_by_embed`,
`skoots.lib.cropper.crops`), imported through oracle/ref_shim.py from /root/reference or — on the GPU box —
from the byte-identical copy under oracle/_ref.  `skoots.lib.eval.eval` itself cannot run (zarr, fastremap,
bism and a checkpoint are absent), so its post-UNet section is replayed around those functions:

    eval.py:223       skeleton = efficient_flood_fill(skeleton)            (int16, in place)
    eval.py:245-284   for crop in crops(vectors, [500,500,50], (50,50,5)): vector_to_embedding(N=10) ;
                      += origin ; index_skeleton_by_embed ; write the interior

`crop=None` is the whole-volume form (the lib functions applied directly, SURVEY §8d C1 rule).
"""
import contextlib
import io
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import ref_shim  # noqa: E402

_ns = None


def available() -> bool:
    return ref_shim.reference_available()


def kind() -> str:
    return ref_shim.reference_kind()


def load():
    global _ns
    if _ns is None:
        ref_shim.install(need_morphology=False)
        from skoots.lib.cropper import crops
        from skoots.lib.flood_fill import efficient_flood_fill
        from skoots.lib.skeleton import index_skeleton_by_embed
        from skoots.lib.vector_to_embedding import vector_to_embedding

        for fn in (crops, efficient_flood_fill, index_skeleton_by_embed, vector_to_embedding):
            # the checker must be the reference itself: never the product's functions left bound by patch_skoots()
            assert getattr(fn, "__module__", "").startswith("skoots."), f"{fn} is not the reference's own function"

        class NS:
            pass
        ns = NS()
        ns.crops, ns.flood, ns.index, ns.v2e = crops, efficient_flood_fill, index_skeleton_by_embed, vector_to_embedding
        _ns = ns
    return _ns


def _quiet(fn, *a, **k):  # the reference prints progress bars
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def flood_fill(skeleton_mask: torch.Tensor) -> torch.Tensor:
    """eval.py:223 on a copy: (X,Y,Z) u8/int16 mask -> int16 labels."""
    ref = load()
    return _quiet(ref.flood, skeleton_mask.to(torch.int16).clone())


def assemble(labels_i16: torch.Tensor, vectors: torch.Tensor, scale: torch.Tensor, N: int, decay: float, crop, overlap,
             out_dtype=torch.int16) -> torch.Tensor:
    ref = load()
    skeleton = labels_i16.unsqueeze(0).unsqueeze(0)
    if crop is None:
        emb = ref.v2e(scale=scale, vector=vectors.unsqueeze(0), N=N, decay=decay)
        return ref.index(skeleton=skeleton, embed=emb)[0, 0].to(out_dtype)
    inst = torch.zeros(labels_i16.shape, dtype=out_dtype)
    size = list(crop)
    with_overlap = all(int(o) > 0 for o in overlap)
    for vec, (x, y, z) in ref.crops(vectors, crop_size=size, overlap=tuple(overlap)):
        emb = ref.v2e(scale=scale, vector=vec, N=N, decay=decay)
        emb += torch.tensor((x, y, z)).view(1, 3, 1, 1, 1)
        got = ref.index(skeleton=skeleton, embed=emb).squeeze()
        if with_overlap:
            inst[x + overlap[0]:x + size[0] - overlap[0], y + overlap[1]:y + size[1] - overlap[1],
                 z + overlap[2]:z + size[2] - overlap[2]] = got[overlap[0]:-overlap[0], overlap[1]:-overlap[1],
                                                                overlap[2]:-overlap[2]]
        else:
            inst[x:x + size[0], y:y + size[1], z:z + size[2]] = got
    return inst


def postprocess(skeleton_mask: torch.Tensor, vectors: torch.Tensor, scale, N: int = 1, decay: float = 1.0, crop=None,
                overlap=(0, 0, 0), out_dtype=torch.int32) -> torch.Tensor:
    """flood fill + assembly: same contract as oracle/skoots_oracle.postprocess, executed by the reference."""
    scale = scale if isinstance(scale, torch.Tensor) else torch.tensor(scale)
    labels = flood_fill(skeleton_mask)
    return assemble(labels, vectors, scale, N, decay, crop, overlap, out_dtype)
