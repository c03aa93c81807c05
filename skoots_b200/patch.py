"""Rebinds the reference's by-name imports to the B200 implementations.

    import skoots_b200.patch; skoots_b200.patch.patch_skoots()

Only modules that are importable are touched; `from x import y` copies are rebound in every already-
imported caller module listed in SURVEY.md §8b.
"""
from __future__ import annotations

import importlib
import sys
from typing import List, Tuple

# (defining module, attribute) -> replacement (module path, attribute)
_TARGETS = {
    ("skoots.lib.vector_to_embedding", "vector_to_embedding"): ("skoots_b200.lib.vector_to_embedding", "vector_to_embedding"),
    ("skoots.lib.skeleton", "index_skeleton_by_embed"): ("skoots_b200.lib.skeleton", "index_skeleton_by_embed"),
    ("skoots.lib.skeleton", "bake_skeleton"): ("skoots_b200.lib.skeleton", "bake_skeleton"),
    ("skoots.lib.skeleton", "skeleton_to_mask"): ("skoots_b200.lib.skeleton", "skeleton_to_mask"),
    ("skoots.lib.skeleton", "average_baked_skeletons"): ("skoots_b200.lib.skeleton", "average_baked_skeletons"),
    ("skoots.lib.flood_fill", "efficient_flood_fill"): ("skoots_b200.lib.flood_fill", "efficient_flood_fill"),
    ("skoots.lib.morphology", "binary_dilation"): ("skoots_b200.lib.morphology", "binary_dilation"),
    ("skoots.lib.morphology", "binary_dilation_2d"): ("skoots_b200.lib.morphology", "binary_dilation_2d"),
    ("skoots.lib.morphology", "binary_erosion"): ("skoots_b200.lib.morphology", "binary_erosion"),
    ("skoots.lib.embedding_to_prob", "baked_embed_to_prob"): ("skoots_b200.lib.embedding_to_prob", "baked_embed_to_prob"),
    # the crop grid (host index arithmetic; same yields, raises where the reference would loop forever — SURVEY B#16)
    ("skoots.lib.cropper", "crops"): ("skoots_b200.lib.cropper", "crops"),
    ("skoots.lib.cropper", "get_total_num_crops"): ("skoots_b200.lib.cropper", "get_total_num_crops"),
    # row f2: the validation metrics (skoots/validate/lib.py:170-275), bound by name in skoots/validate/__main__.py:9-13
    ("skoots.validate.lib", "mask_iou"): ("skoots_b200.validate", "mask_iou"),
    ("skoots.validate.lib", "mask_dice"): ("skoots_b200.validate", "mask_dice"),
    ("skoots.validate.lib", "accuracies_from_iou"): ("skoots_b200.validate", "accuracies_from_iou"),
}

# modules holding `from ... import name` copies (SURVEY.md §8b "bound at")
_CALLERS = (
    "skoots.lib.eval", "skoots.lib.flood_fill", "skoots.train.engine", "skoots.train.merged_transform", "skoots.train.loss",
    "skoots.experimental.sparse_engine", "skoots.experimental.eval", "skoots.experimental.sparse_loss",
    "skoots.experimental.sparse_transforms", "skoots.experimental.modifiers", "skoots.validate.__main__",
)


_SAVED: dict = {}


def unpatch_skoots() -> None:
    """Restores every attribute patch_skoots() rebound."""
    for (mod_name, attr), old in list(_SAVED.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, old)
    _SAVED.clear()


def _bug_compatible_flood_fill():
    """`efficient_flood_fill` with the reference's multi-crop behaviour (seam heuristic, label re-use after an empty crop:
    SURVEY.md B#6-#8) reproduced bit for bit on volumes larger than one 1000x1000x200 crop."""
    import functools

    from skoots_b200.lib.flood_fill import efficient_flood_fill
    bound = functools.partial(efficient_flood_fill, reference_crops=True)
    functools.update_wrapper(bound, efficient_flood_fill)
    return bound


def _bug_compatible_bake_skeleton():
    """`bake_skeleton` with the reference's own dispatch: the semantics of its Triton kernel for CUDA masks
    (skeleton.py:505-512), of its CPU path for host tensors."""
    import functools

    from skoots_b200.lib.skeleton import bake_skeleton
    bound = functools.partial(bake_skeleton, triton_compat="auto")
    functools.update_wrapper(bound, bake_skeleton)
    return bound


def patch_skoots(bug_compatible: bool = False) -> List[Tuple[str, str]]:
    """Returns the (module, attribute) pairs that were rebound. Idempotent.

    bug_compatible=False (default) binds the exact connected-component labelling: identical to the reference on any
    volume that fits one of its 1000x1000x200 flood-fill crops, and the same partition minus the reference's spurious
    seam merges on larger ones.  bug_compatible=True binds `efficient_flood_fill(..., reference_crops=True)` instead:
    the reference's own result bit for bit on every volume, quirks included (north_star's "bit-exact to the reference"),
    and `bake_skeleton(..., triton_compat="auto")`: CUDA masks get what the reference's Triton kernel returns (fp16,
    anisotropy on the squared differences, per-axis maximum on ties, phantom origin point), host masks its CPU result."""
    done: List[Tuple[str, str]] = []
    originals = {}
    special = {("skoots.lib.flood_fill", "efficient_flood_fill"): _bug_compatible_flood_fill(),
               ("skoots.lib.skeleton", "bake_skeleton"): _bug_compatible_bake_skeleton()} if bug_compatible else {}
    for (mod_name, attr), (new_mod, new_attr) in _TARGETS.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            continue
        new = special.get((mod_name, attr)) or getattr(importlib.import_module(new_mod), new_attr)
        old = getattr(mod, attr, None)
        if old is not None and old is not new:
            originals[id(old)] = new
            _SAVED.setdefault((mod_name, attr), old)
        setattr(mod, attr, new)
        done.append((mod_name, attr))
    by_name = {attr: special.get(key) or getattr(importlib.import_module(nm), na) for key, (nm, na) in _TARGETS.items() for attr in (key[1],)}
    for caller in _CALLERS:
        mod = sys.modules.get(caller)
        if mod is None:
            continue
        for attr, new in by_name.items():
            cur = getattr(mod, attr, None)
            if cur is not None and cur is not new and (id(cur) in originals or getattr(cur, "__module__", "").startswith(("skoots.lib", "skoots.validate"))):
                _SAVED.setdefault((caller, attr), cur)
                setattr(mod, attr, new)
                done.append((caller, attr))
    return done
