"""TEST INFRASTRUCTURE — import shim that lets the *unmodified* reference package
(`/root/reference/skoots`) load in the build container, where `skimage`, `bism` and
`yacs` are not installed.  Used only by `oracle/gen_golden.py` and by the
`tests/test_oracle_vs_reference.py` pin (skipped when /root/reference is absent,
e.g. on the GPU box).  Nothing under `skoots_b200/` imports this file.

The three stubs cover exactly what the reference touches at import time:
  * `skimage.morphology.disk`   (skoots/lib/utils.py:6-14, used at utils.py:423-424)
  * `skimage.io`                (skoots/validate/utils.py:4; only so that skoots.validate.lib imports)
  * `bism.*`                    (skoots/lib/utils.py:6-14 model factory imports)
  * `yacs.config.CfgNode`       (type annotation only)
Used also by `oracle/ref_runner.py` (bench.py's CPU baseline).
`PYTORCH_JIT=0` must be set before torch is imported: the scripted morphology
functions hash a list through functools.cache and fail under torch 2.11
(skoots/lib/morphology.py:10-16,145,167).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """/root/reference in the build container; on the GPU box the byte-identical copy that
    oracle/build_ref.py placed under oracle/_ref (git-ignored, travels with the snapshot)."""
    env = os.environ.get("SKOOTS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "skoots", "lib")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "skoots", "lib"))


def reference_kind() -> str:
    return "live tree" if REFERENCE_ROOT == "/root/reference" else "oracle/_ref copy"


def _disk(radius, dtype=None):
    import numpy as np

    span = np.arange(-radius, radius + 1)
    xx, yy = np.meshgrid(span, span)
    return np.asarray((xx * xx + yy * yy) <= radius * radius, dtype=dtype or np.uint8)


class _Fabricating(types.ModuleType):
    """module whose attributes are fabricated sub-modules (for `from bism.x.y import Z`)."""

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        child = _Fabricating(self.__name__ + "." + key)
        sys.modules[child.__name__] = child
        setattr(self, key, child)
        return child


def install(need_morphology: bool = True):
    """Make `import skoots.lib.*` work. Idempotent. Raises if the reference is absent.
    need_morphology=False skips the PYTORCH_JIT=0 requirement: only the scripted morphology functions
    fail under torch 2.11, and the assembly path bench.py times (flood fill, vector_to_embedding,
    index_skeleton_by_embed) then runs with the reference's stock torch.jit.script decorators active."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if need_morphology and os.environ.get("PYTORCH_JIT", "1") != "0":
        if "torch" in sys.modules:
            raise RuntimeError("set PYTORCH_JIT=0 before importing torch to run the reference")
        os.environ["PYTORCH_JIT"] = "0"

    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        skm = types.ModuleType("skimage.morphology")
        skm.disk = _disk
        sk.morphology = skm
        ski = types.ModuleType("skimage.io")  # skoots/validate/utils.py:4 imports it; nothing here reads a file

        def _no_io(*a, **k):
            raise RuntimeError("skimage.io is stubbed: the oracle never touches image files")
        ski.imread = ski.imsave = _no_io
        sk.io = ski
        sys.modules["skimage"] = sk
        sys.modules["skimage.morphology"] = skm
        sys.modules["skimage.io"] = ski
    if "bism" not in sys.modules:
        for name in ("bism", "bism.backends", "bism.modules", "bism.models",
                     "bism.models.spatial_embedding"):
            sys.modules[name] = _Fabricating(name)
        sys.modules["bism.models.spatial_embedding"].SpatialEmbedding = object
    if "yacs" not in sys.modules:
        y = types.ModuleType("yacs")
        yc = types.ModuleType("yacs.config")
        yc.CfgNode = dict
        y.config = yc
        sys.modules["yacs"] = y
        sys.modules["yacs.config"] = yc
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import skoots.lib.cropper  # noqa: F401
    import skoots.lib.embedding_to_prob  # noqa: F401
    import skoots.lib.flood_fill  # noqa: F401
    import skoots.lib.morphology  # noqa: F401
    import skoots.lib.skeleton  # noqa: F401
    import skoots.lib.utils  # noqa: F401
    import skoots.lib.vector_to_embedding  # noqa: F401
    import skoots

    return skoots
