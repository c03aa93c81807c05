"""ctypes binding of libskoots_b200.so (include/skoots_b200.h).

There is deliberately no fallback: if the shared library or a CUDA device is missing, the call
raises.  PyTorch is used only to own device memory and streams.  The `skoots.lib` mirrors accept
HOST tensors too — the reference's `eval()` and `skoots-validate` hand CPU tensors to these
functions (skoots/lib/eval.py:223,258-284; skoots/validate/__main__.py) — by staging them through
the current CUDA device (`compute_device` / `stage_in` / `stage_out`): H2D, the same kernels, D2H.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Sequence

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libskoots_b200.so")

SKB_U8, SKB_I16, SKB_I32, SKB_F16, SKB_BF16, SKB_F32 = range(6)
STATUS_ROOT_OVERFLOW = 1
STATUS_MISSING_ID = 2
STATUS_PEER_TIMEOUT = 4
STATUS_LABEL_RANGE = 8
STATUS_HALO_RANGE = 16
PEER_HANDLE_BYTES = 64
MAX_WORLD = 16
CCL_WORKSPACE_CLEAN = 1
CCL_PHASE_PACK = 2
CCL_PHASE_LABEL = 4
CCL_TILES_DYNAMIC = 8
CCL_TILES_STATIC = 16
CCL_KEEP_STATUS = 32

_DTYPES = {
    torch.uint8: SKB_U8, torch.bool: SKB_U8, torch.int16: SKB_I16, torch.int32: SKB_I32,
    torch.float16: SKB_F16, torch.bfloat16: SKB_BF16, torch.float32: SKB_F32,
}

_lib = None

_c_i64, _c_int, _c_vp, _c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
_c_f3 = ctypes.POINTER(ctypes.c_float)
_c_i3 = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); must list every symbol include/skoots_b200.h declares
SIGNATURES = {
    "skb_version": (_c_int, []),
    "skb_last_error": (ctypes.c_char_p, []),
    "skb_vec_embed3d": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_f3, _c_int, ctypes.c_double, _c_vp, _c_vp]),
    "skb_vec_embed2d": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_vp]),
    "skb_vec_embed_bwd": (_c_int, [_c_vp, _c_i64, _c_int, _c_i64, _c_f3, _c_vp, _c_int, _c_vp]),
    "skb_index_by_embed": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp]),
    "skb_ccl_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_i64, _c_i64]),
    "skb_ccl_label_sparse": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_int, ctypes.c_int32, _c_i64, _c_vp, _c_sz, _c_vp, _c_vp, _c_int, _c_vp]),
    "skb_ccl_write_dense": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_int, _c_vp]),
    "skb_stencil3": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_int, _c_vp]),
    "skb_masked_mean27": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp]),
    "skb_tile_epilogue": (_c_int, [_c_vp, _c_int, _c_int, _c_i3, _c_i3, _c_i3, ctypes.c_float, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp]),
    "skb_embed_prob_fwd": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_i64, _c_f3, ctypes.c_float, _c_vp, _c_vp]),
    "skb_embed_prob_bwd": (_c_int, [_c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_i64, _c_int, _c_i64, _c_f3, ctypes.c_float, _c_vp, _c_vp, _c_vp]),
    "skb_vec_prob": (_c_int, [_c_vp, _c_int, _c_vp, _c_int, _c_i64, _c_int, _c_i64, _c_i64, _c_i64, _c_f3, _c_f3, ctypes.c_float, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_bake_skeletons": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_f3, _c_int,
                                    _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_stamp_disks": (_c_int, [_c_vp, _c_int, _c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_shard_label_local": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_sz, _c_vp, _c_int, _c_vp]),
    "skb_shard_emit_runs": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_int, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp]),
    "skb_shard_ingest_runs": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_shard_clear_halo": (_c_int, [_c_i64, _c_vp, _c_i64, _c_vp, _c_vp]),
    "skb_shard_clear_halo_peer": (_c_int, [_c_i64, _c_vp, _c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_shard_boundary_pairs": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_shard_merge": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_int, _c_int, _c_i64, _c_i64, ctypes.c_int32, _c_vp, _c_vp, _c_vp]),
    "skb_assemble_stream": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp]),
    "skb_assemble_resolve": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp]),
    "skb_label_max": (_c_int, [_c_vp, _c_int, _c_i64, _c_vp, _c_vp]),
    "skb_renumber_workspace_bytes": (_c_sz, [_c_i64, _c_i64]),
    "skb_renumber": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_sz, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_apply_label_table": (_c_int, [_c_vp, _c_int, _c_i64, _c_vp, _c_i64, _c_vp]),
    "skb_unique_index_workspace_bytes": (_c_sz, [_c_i64]),
    "skb_unique_index": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_vp, _c_vp]),
    "skb_contingency": (_c_int, [_c_vp, _c_int, _c_vp, _c_int, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_iou_dice": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp]),
    "skb_accuracies_from_iou": (_c_int, [_c_vp, _c_i64, _c_i64, ctypes.c_float, _c_vp, _c_vp, _c_vp]),
    "skb_elastic_resample": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp]),
    "skb_elastic_points": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_peer_alloc": (_c_int, [_c_sz, ctypes.POINTER(_c_vp)]),
    "skb_peer_free": (_c_int, [_c_vp]),
    "skb_peer_export": (_c_int, [_c_vp, ctypes.c_char_p]),
    "skb_peer_open": (_c_int, [ctypes.c_char_p, ctypes.POINTER(_c_vp)]),
    "skb_peer_close": (_c_int, [_c_vp]),
    "skb_shard_mailbox_bytes": (_c_sz, [_c_int, _c_i64, _c_i64, _c_i64]),
    "skb_shard_begin": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_vp]),
    "skb_shard_emit_runs_peer": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_shard_ingest_runs_peer": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_shard_push": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, ctypes.POINTER(ctypes.c_uint64), _c_int, _c_int, _c_i64, _c_i64, _c_i64,
                                _c_vp, _c_vp]),
    "skb_shard_begin_pass": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "skb_shard_merge_peer": (_c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_int, _c_int, _c_i64, _c_i64, _c_i64, ctypes.c_int32, _c_vp, _c_vp, _c_vp]),
    "skb_assemble_slab": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp]),
    "skb_assemble_slab_ex": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_f3, _c_int, ctypes.c_double, _c_i3, _c_i3,
                                      _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp]),
    "skb_assemble_planar": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_f3, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_vp]),
    "skb_assemble_range": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_f3, _c_int, ctypes.c_double, _c_i3, _c_i3, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_i64, _c_i64, _c_vp]),
    "skb_assemble": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_f3, _c_int, ctypes.c_double, _c_i3, _c_i3, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_vp]),
}


class SkootsB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkootsB200Error(
                f"{LIB_PATH} is missing: build it with `python -m skoots_b200.build` "
                "(skoots_b200 has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise SkootsB200Error(f"libskoots_b200 error {rc}: {load().skb_last_error().decode()}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise SkootsB200Error(f"unsupported dtype {t.dtype}") from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise SkootsB200Error("skoots_b200 operates on CUDA tensors only (no CPU fallback)")
        if dev is not None and t.device != dev:
            raise SkootsB200Error("all tensors must live on the same CUDA device")
        dev = t.device
    return dev


def compute_device(*tensors: torch.Tensor):
    """(device the kernels run on, staged).  All-CUDA arguments run where they live; all-HOST arguments are staged
    through the current CUDA device (staged = True: the caller copies inputs up and results back); a mix is an
    error, as it is in the reference's torch code.  Without a CUDA device this raises: nothing here computes on the CPU."""
    cuda = [t for t in tensors if isinstance(t, torch.Tensor) and t.is_cuda]
    if len(cuda) == len(tensors):
        return require_cuda(*tensors), False
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise SkootsB200Error("skoots_b200 expects torch tensors")
    if cuda:
        raise SkootsB200Error("all tensors must live on the same device (got a mix of host and CUDA tensors)")
    if not torch.cuda.is_available():
        raise SkootsB200Error("no CUDA device: skoots_b200 stages host tensors through the GPU and has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()), True


_STAGE_CACHE: dict = {}


def stage_in(t: torch.Tensor, dev: torch.device, reuse: bool = False) -> torch.Tensor:
    """host -> device copy of a staged argument.  reuse=True keeps the device copy of the LAST such tensor (keyed by
    storage, shape, dtype and in-place version) so that a caller that passes the same large host tensor call after
    call — the whole-image label volume in eval()'s crop loop, skoots/lib/eval.py:277-279 — uploads it once."""
    if t.device == dev:
        return t
    if not reuse:
        return t.to(dev)
    import weakref
    key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version, str(dev))
    owner = t if t._base is None else t._base  # views share the base's storage and version counter
    hit = _STAGE_CACHE.get("labels")
    if hit is not None and hit[0] == key and hit[1]() is owner:
        return hit[2]
    d = t.to(dev)
    _STAGE_CACHE["labels"] = (key, weakref.ref(owner), d)
    return d


def stage_out(t, staged: bool):
    """device -> host copy of a staged call's result (tensors inside tuples included)."""
    if not staged:
        return t
    if isinstance(t, tuple):
        return tuple(stage_out(v, True) for v in t)
    return t.cpu() if isinstance(t, torch.Tensor) else t


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f3(values: Sequence[float]):
    vals = [float(v) for v in values]
    return (ctypes.c_float * len(vals))(*vals)


def i3(values: Sequence[int]):
    vals = [int(v) for v in values]
    return (ctypes.c_int32 * len(vals))(*vals)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _nvml_handle(device):
    """NVML handle of a torch CUDA device (logical index -> physical GPU through its UUID or PCI address)."""
    import pynvml
    pynvml.nvmlInit()
    props = torch.cuda.get_device_properties(device)
    try:
        return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
    except Exception:
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())


@contextlib.contextmanager
def numa_local(device):
    """While active, the calling thread runs only on the CPUs that NVML reports as local to `device`, so that pinned host
    memory allocated inside (cudaHostAlloc places its pages at allocation time, on the node of the allocating thread) lands
    on the GPU's own NUMA node and its DMA does not cross the socket interconnect — what limits the uploads when all
    eight GPUs of a box copy at once (profiles/r02_h2d_probe_g8.json).  Best effort: without NVML, a topology or the
    permission it changes nothing.  Yields a dict describing what happened; the previous affinity is restored on exit."""
    info = {"bound": False}
    previous = None
    try:
        pynvml, handle = _nvml_handle(device)
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(allowed) + 64) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        try:
            info["node"] = int(pynvml.nvmlDeviceGetNumaNodeId(handle))
        except Exception:
            pass
        info["local_cpus"], info["allowed_cpus"] = len(local), len(allowed)
        if local and local != allowed:
            os.sched_setaffinity(0, local)
            previous = allowed
            info["bound"] = True
    except Exception as exc:  # no NVML / no topology / not permitted: the allocation simply stays where the OS puts it
        info["why"] = repr(exc)[:120]
    try:
        yield info
    finally:
        if previous is not None:
            os.sched_setaffinity(0, previous)
