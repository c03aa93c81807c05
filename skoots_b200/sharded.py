"""Z-sharded post-processing across the GPUs of one box.

One process per GPU; rank r owns the slab z in [z0, z1) of the u8 skeleton mask and of the fp16 vector
field.  A pass has two exchange steps, both tiny compared with the slab itself:

  1. neighbour send/recv (NCCL P2P): the z-runs in my first / last H planes with their component ids
     -> the neighbour's halo (H = ceil(scale_z), the farthest a vector can point along z);
  2. all-gather: every rank's component roots and the (my root, neighbour root) pairs found on the
     slab faces -> every rank solves the same union-find and numbers all components identically
     (the single-GPU numbering).

Everything else is the single-GPU kernels restricted to the slab.  Two transports carry the exchanges:

  * "peer" (default on a multi-GPU box, `PeerComm`): every rank owns a mailbox in peer-visible device
    memory; the producing kernels store straight into the consumer GPU's mailbox over NVLink and
    release a flag there, the consuming kernels spin on flags in their own HBM.  No NCCL call and no
    host involvement inside a pass: a pass is ~30 kernel launches on one stream = one CUDA graph.
    torch.distributed is used once, at set-up, to swap the 64-byte IPC handles of the mailboxes.
  * "nccl" (`TorchDistComm`; gloo in the CPU tests of the plumbing): NCCL send/recv + all-gather of
    fixed-size buffers between the phases.

`LocalGroup` runs all ranks of either transport inside one process on one GPU (the GPU parity test
checks the sharded result is bit-identical to the unsharded one).
"""
from __future__ import annotations

import contextlib
import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L


def slab_bounds(Z: int, world: int) -> List[Tuple[int, int]]:
    """contiguous z ranges, multiples of 64 planes (the bit-mask word), as even as possible."""
    if Z % 64 != 0:
        raise ValueError("sharded post-processing needs Z to be a multiple of 64")
    words = Z // 64
    if words < world:
        raise ValueError(f"Z={Z} has only {words} 64-plane words: cannot shard over {world} ranks")
    cuts = [(words * r) // world * 64 for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange_layout(cap_roots: int, cap_pairs: int) -> int:
    """int32 elements of one rank's all-gather payload: [n_roots, n_pairs, roots, pairs]."""
    return 2 + cap_roots + 2 * cap_pairs


class TorchDistComm:
    """collectives over torch.distributed (NCCL on GPUs; gloo in the CPU tests of the plumbing)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def neighbour_exchange(self, send_lo: Tensor, send_hi: Tensor, recv_lo: Tensor, recv_hi: Tensor) -> None:
        """send_lo -> rank-1 (arrives in its recv_hi); send_hi -> rank+1 (arrives in its recv_lo)."""
        d, ops = self.dist, []
        if self.rank > 0:
            ops.append(d.P2POp(d.isend, send_lo, self.rank - 1, self.group))
            ops.append(d.P2POp(d.irecv, recv_lo, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(d.P2POp(d.isend, send_hi, self.rank + 1, self.group))
            ops.append(d.P2POp(d.irecv, recv_hi, self.rank + 1, self.group))
        if ops:
            for req in d.batch_isend_irecv(ops):
                req.wait()

    def all_gather(self, out: Tensor, src: Tensor) -> None:
        self.dist.all_gather_into_tensor(out, src, group=self.group)

    def barrier(self) -> None:
        self.dist.barrier(self.group)


class Mailbox:
    """One rank's mailbox: cudaMalloc'ed by the library (IPC-exportable), freed with the object."""

    def __init__(self, lib, nbytes: int, device):
        self.lib, self.nbytes, self.device = lib, int(nbytes), torch.device(device)
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            L.check(lib.skb_peer_alloc(self.nbytes, ctypes.byref(out)))
        self.ptr = int(out.value)

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(L.PEER_HANDLE_BYTES)
        L.check(self.lib.skb_peer_export(self.ptr, buf))
        return buf.raw

    def ints(self, byte_offset: int, count: int) -> Tensor:
        """a torch view of part of the mailbox (diagnostics / tests only)."""
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (self.ptr + byte_offset, False),
                                      "version": 2, "strides": None}
        return torch.as_tensor(v, device=self.device)

    def free(self) -> None:
        if self.ptr:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                self.lib.skb_peer_free(self.ptr)
            self.ptr = 0

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


def _device_view(ptr: int, shape, dtype: torch.dtype, device) -> Tensor:
    """a torch tensor over library-allocated device memory (CUDA array interface; the memory outlives the view's users)"""
    typestr = {torch.float16: "<f2", torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4", torch.bfloat16: "<i2"}[dtype]

    class _View:
        pass
    v = _View()
    v.__cuda_array_interface__ = {"shape": tuple(int(n) for n in shape), "typestr": typestr, "data": (int(ptr), False),
                                  "version": 2, "strides": None}
    t = torch.as_tensor(v, device=device)
    return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t


class PeerComm:
    """Peer transport across processes: allocates this rank's mailbox and maps every other rank's through
    CUDA IPC (handles swapped once over torch.distributed).  After `connect` no collective is issued."""

    transport = "peer"

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.mailbox: Optional[Mailbox] = None
        self.peer_ptrs: List[int] = []
        self._opened: List[int] = []

    def connect(self, lib, nbytes: int, device) -> Tuple[Mailbox, List[int]]:
        """Collective.  Either every rank ends up with every mailbox mapped, or every rank raises
        SkootsB200Error (a rank that cannot allocate / export / map must not leave the others waiting)."""
        self.lib = lib
        error = ""
        try:
            self.mailbox = Mailbox(lib, nbytes, device)
            handle = self.mailbox.handle()
        except L.SkootsB200Error as exc:
            self.mailbox, handle, error = None, bytes(L.PEER_HANDLE_BYTES), str(exc)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(device)
        every = torch.empty(self.world * L.PEER_HANDLE_BYTES, dtype=torch.uint8, device=device)
        self.dist.all_gather_into_tensor(every, mine, group=self.group)
        handles = every.cpu().numpy().tobytes()
        self.peer_ptrs = []
        if not error:
            with torch.cuda.device(device):
                for r in range(self.world):
                    if r == self.rank:
                        self.peer_ptrs.append(self.mailbox.ptr)
                        continue
                    out = ctypes.c_void_p()
                    h = handles[r * L.PEER_HANDLE_BYTES:(r + 1) * L.PEER_HANDLE_BYTES]
                    rc = lib.skb_peer_open(h, ctypes.byref(out)) if any(h) else -3
                    if rc != 0:
                        error = f"cannot map rank {r}'s mailbox: {lib.skb_last_error().decode()}"
                        break
                    self.peer_ptrs.append(int(out.value))
                    self._opened.append(int(out.value))
        ok = torch.tensor([0 if error else 1], dtype=torch.int32, device=device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)  # also: everything is mapped before anyone stores
        if int(ok.item()) == 0:
            self._release()
            raise L.SkootsB200Error("peer mailboxes are not available on this box"
                                    + (f" ({error})" if error else " (another rank failed)"))
        return self.mailbox, self.peer_ptrs

    def map_neighbours(self, box: "Mailbox", device) -> Tuple[int, int]:
        """Collective.  Shares one more peer-visible buffer of every rank (its slab of the vector field) and maps the two
        Z-neighbours' buffers into this process: (pointer to rank-1's buffer or 0, pointer to rank+1's buffer or 0)."""
        mine = torch.frombuffer(bytearray(box.handle()), dtype=torch.uint8).to(device)
        every = torch.empty(self.world * L.PEER_HANDLE_BYTES, dtype=torch.uint8, device=device)
        self.dist.all_gather_into_tensor(every, mine, group=self.group)
        handles = every.cpu().numpy().tobytes()
        ptrs, error = [0, 0], ""
        with torch.cuda.device(device):
            for k, r in enumerate((self.rank - 1, self.rank + 1)):
                if r < 0 or r >= self.world:
                    continue
                out = ctypes.c_void_p()
                rc = self.lib.skb_peer_open(handles[r * L.PEER_HANDLE_BYTES:(r + 1) * L.PEER_HANDLE_BYTES], ctypes.byref(out))
                if rc != 0:
                    error = f"cannot map rank {r}'s vector slab: {self.lib.skb_last_error().decode()}"
                    break
                ptrs[k] = int(out.value)
                self._opened.append(int(out.value))
        ok = torch.tensor([0 if error else 1], dtype=torch.int32, device=device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            raise L.SkootsB200Error("the neighbours' vector slabs cannot be mapped" + (f" ({error})" if error else " (another rank failed)"))
        self._boxes = getattr(self, "_boxes", []) + [box]
        return ptrs[0], ptrs[1]

    def _release(self) -> None:
        for box in getattr(self, "_boxes", []):
            box.free()
        self._boxes = []
        if self._opened:
            with torch.cuda.device(self.mailbox.device if self.mailbox else torch.cuda.current_device()):
                for p in self._opened:
                    self.lib.skb_peer_close(p)
        self._opened = []
        self.peer_ptrs = []
        if self.mailbox is not None:
            self.mailbox.free()
            self.mailbox = None

    def barrier(self) -> None:
        self.dist.barrier(self.group)

    def close(self) -> None:
        """unmap the peers' mailboxes, then (after a barrier: nobody may still be storing into mine) free mine."""
        if self.mailbox is None:
            return
        torch.cuda.synchronize(self.mailbox.device)
        with torch.cuda.device(self.mailbox.device):
            for p in self._opened:
                self.lib.skb_peer_close(p)
        self._opened = []
        self.dist.barrier(self.group)
        self._release()


class ShardedAssembler:
    """Per-rank state + the phases of one pass.

    N = 1 (the headline mode) needs only the label halo.  With N > 1 — `eval()` runs N = 10 over its 500x500x50 crop
    grid, skoots/lib/eval.py:245-284 — a walk stays inside the crop that owns its voxel, so its hops can leave the slab by
    up to crop_z - overlap_z - 1 planes (44 for eval()): every pass starts by swapping that many planes of the vector
    field with the two Z-neighbours ("vector halo"; one NCCL send/recv pair per face, or plain copies in
    the single-process emulation), and the label halo defaults to a whole 64-plane word because a ten-hop walk is not
    bounded by scale_z."""

    def __init__(self, shape: Sequence[int], world: int, rank: int, device, scale=(60, 60, 12), hops: int = 1,
                 decay: float = 1.0, crop: Optional[Sequence[int]] = None, overlap: Sequence[int] = (0, 0, 0), comm=None,
                 halo: Optional[int] = None, cap_roots: int = 1 << 18, cap_pairs: int = 1 << 17, cap_runs: Optional[int] = None,
                 out_dtype=torch.int32, split: Optional[bool] = None):
        X, Y, Z = (int(v) for v in shape)
        self.hops, self.decay = int(hops), float(decay)
        self._crop = L.i3(crop) if crop is not None else None
        self._overlap = L.i3(overlap) if crop is not None else None
        self.vhalo_lo = self.vhalo_hi = None
        self.vh = 0
        self.shape, self.world, self.rank, self.dev = (X, Y, Z), world, rank, torch.device(device)
        self.scale = [float(s) for s in scale]
        bounds = slab_bounds(Z, world)
        self.z_range = bounds[rank]
        self.Zl = self.z_range[1] - self.z_range[0]
        if self.hops > 1:
            if split:
                raise ValueError("the stream/resolve split covers N = 1 only")
            cz = min(int(crop[2]), Z) if crop is not None else Z
            ov_z = int(overlap[2]) if crop is not None else 0
            self.vh = cz - ov_z - 1 if ov_z > 0 else cz - 1   # how far a hop can be from its voxel along z (it stays in the owner crop)
            if world > 1 and self.vh > min(b[1] - b[0] for b in bounds):
                raise ValueError(f"N > 1 over crops {cz} planes deep needs {self.vh} planes of the neighbour's vector field, "
                                 f"more than the thinnest slab holds: use fewer ranks or the whole-volume pass")
            if halo is None:
                halo = min(64, self.Zl)
        # label halo: how far beyond a face this rank can answer a gather target.  The network's vectors lie in [-1, 1]
        # (vector_to_embedding.py:140), so ceil(scale_z) planes is the default; a field that exceeds it is DETECTED
        # (STATUS_HALO_RANGE -> check_status() raises) and the caller can ask for up to 64 planes
        self.halo = int(math.ceil(abs(self.scale[2]))) if halo is None else int(halo)
        if self.halo > 64 or self.halo > self.Zl:
            raise ValueError("halo deeper than one 64-plane word / than the slab is not supported")
        self.comm = comm
        self.transport = getattr(comm, "transport", "nccl")
        self.lib = L.load()
        self.capacity = max(1 << 16, (X * Y * self.Zl) // 8)
        need = self.lib.skb_ccl_workspace_bytes(X, Y, Z, self.capacity)
        self.workspace = torch.empty(need, dtype=torch.uint8, device=self.dev)
        # boundary runs per face: the whole (fixed-size) buffer travels, so keep it close to what a skeleton mask
        # needs (~0.1 % of the face voxels start a run); overflow is reported through the status word
        self.cap_runs = int(cap_runs) if cap_runs else max(1 << 14, (X * Y * self.halo) // 128)
        self.cap_roots, self.cap_pairs = cap_roots, cap_pairs
        mk = lambda n, dt=torch.int32: torch.zeros(n, dtype=dt, device=self.dev)
        self.halo_lo = mk(X * Y, torch.int64) if rank > 0 else None
        self.halo_hi = mk(X * Y, torch.int64) if rank < world - 1 else None
        self.exch = mk(exchange_layout(cap_roots, cap_pairs))
        self.mailbox: Optional[Mailbox] = None
        self.peer_ptrs: List[int] = []
        if self.transport == "peer":
            if world > L.MAX_WORLD:
                raise ValueError(f"the peer transport supports up to {L.MAX_WORLD} ranks")
            self._geom = (world, self.cap_runs, cap_roots, cap_pairs)
            nbytes = self.lib.skb_shard_mailbox_bytes(*self._geom)
            if nbytes == 0:
                raise L.SkootsB200Error(self.lib.skb_last_error().decode())
            self.mailbox_bytes = int(nbytes)
            if comm is not None and hasattr(comm, "connect"):
                self.attach(*comm.connect(self.lib, self.mailbox_bytes, self.dev))
        else:
            self.send_lo, self.send_hi = mk(3 * (self.cap_runs + 1)), mk(3 * (self.cap_runs + 1))
            self.recv_lo, self.recv_hi = mk(3 * (self.cap_runs + 1)), mk(3 * (self.cap_runs + 1))
            self.gathered = mk(world * exchange_layout(cap_roots, cap_pairs))
        self.meta = mk(2)  # [n_components, status]
        self.out = torch.empty((X, Y, self.Zl), dtype=out_dtype, device=self.dev)
        self._vec_dtype = None  # vector halos are allocated when the field's dtype is known (load / run_host)
        # N > 1 over the peer transport: the slab of the vector field lives in peer-visible memory and the two Z-neighbours'
        # slabs are mapped into this process; the few hops that cross a face read the neighbour GPU's memory over NVLink
        # directly — no vector planes are exchanged (the NCCL transport swaps packed copies of the faces every pass)
        self.peer_vec = self.transport == "peer" and self.hops > 1 and world > 1
        self._vec_box: Optional[Mailbox] = None
        self.vdepth_lo = self.vdepth_hi = 0
        # stream/resolve split of the gather (pipeline.assemble_split): the slab's stream phase runs on the current
        # stream while the labelling chain AND both exchanges run on a high-priority side stream
        can_split = (X * Y * self.Zl) % 256 == 0
        if split and not can_split:
            raise ValueError("the stream/resolve split needs the slab to be a multiple of 256 voxels")
        self.split = bool(split)  # off by default: measured slower than the fused slab gather
        self.flags = torch.empty(X * Y * self.Zl // 256, dtype=torch.int32, device=self.dev) if self.split else None
        from .pipeline import stream_ctas_default
        self.stream_ctas = stream_ctas_default()
        self._clean = False
        self.mask: Optional[Tensor] = None
        self.vec: Optional[Tensor] = None
        # kernel launches of one pass on this rank (bench.py reports them as gpu_launches)
        faces = (rank > 0) + (rank < world - 1)
        if self.transport == "peer":
            # begin=1, clear_halos=1, local(init,pack,tile,boundary,roots)=5, emit+signal=2, ingest=faces,
            # push (signals itself)=1, merge(init,union,mark,scan x2,rank,publish x2)=8, gather=1
            self.launches_per_step = 1 + (faces > 0) + 5 + 2 * (faces > 0) + faces + 1 + 8 + 1
        else:
            # clear_halo=faces, local=5, emit=faces, pack_roots=1, ingest=faces, merge=8, gather=1 (+ NCCL's own kernels)
            self.launches_per_step = faces + 5 + faces + 1 + faces + 8 + 1
        self.launches_per_step += 1 if self.split else 0  # split: stream + resolve instead of one gather

    def attach(self, mailbox: Mailbox, peer_ptrs: Sequence[int]) -> None:
        """peer transport: my mailbox and the (mapped) base pointers of every rank's mailbox, own included."""
        assert len(peer_ptrs) == self.world and peer_ptrs[self.rank] == mailbox.ptr
        self.mailbox, self.peer_ptrs = mailbox, [int(p) for p in peer_ptrs]
        self._peer_arr = (ctypes.c_uint64 * self.world)(*self.peer_ptrs)

    # ---- data ------------------------------------------------------------------------------------
    def load(self, mask_slab: Tensor, vec_slab: Tensor) -> None:
        X, Y, _ = self.shape
        assert tuple(mask_slab.shape) == (X, Y, self.Zl) and tuple(vec_slab.shape) == (3, X, Y, self.Zl)
        L.require_cuda(mask_slab, vec_slab)
        self.mask = (mask_slab.view(torch.uint8) if mask_slab.dtype == torch.bool else mask_slab).contiguous()
        if self.peer_vec:
            self._peer_vector_slab(vec_slab.dtype)
            self.vec.copy_(vec_slab)
            if self.comm is not None and hasattr(self.comm, "barrier"):
                self.comm.barrier()  # every rank's slab is in place before anybody's walk reads it
        else:
            self.vec = vec_slab.contiguous()
            self._alloc_vector_halos(self.vec.dtype)

    def _peer_vector_slab(self, dtype) -> None:
        """allocates this rank's vector slab in peer-visible memory (once per dtype) and maps the neighbours' slabs."""
        if self._vec_box is not None and self._vec_dtype == dtype:
            return
        X, Y, _ = self.shape
        self._vec_dtype = dtype
        nbytes = 3 * X * Y * self.Zl * torch.empty(0, dtype=dtype).element_size()
        self._vec_box = Mailbox(self.lib, nbytes, self.dev)
        self.vec = _device_view(self._vec_box.ptr, (3, X, Y, self.Zl), dtype, self.dev)
        self.graph = None
        if self.comm is not None and hasattr(self.comm, "map_neighbours"):
            lo, hi = self.comm.map_neighbours(self._vec_box, self.dev)
            bounds = slab_bounds(self.shape[2], self.world)
            self.attach_vector_peers(lo, bounds[self.rank - 1][1] - bounds[self.rank - 1][0] if self.rank > 0 else 0,
                                     hi, bounds[self.rank + 1][1] - bounds[self.rank + 1][0] if self.rank < self.world - 1 else 0)

    def attach_vector_peers(self, lo_ptr: int, lo_depth: int, hi_ptr: int, hi_depth: int) -> None:
        """peer-vector mode: device pointers to the lower / upper neighbour's (3,X,Y,depth) vector slab (0 = none)."""
        self.vhalo_lo, self.vhalo_hi = int(lo_ptr), int(hi_ptr)
        self.vdepth_lo, self.vdepth_hi = int(lo_depth), int(hi_depth)

    def _alloc_vector_halos(self, dtype) -> None:
        if self.hops == 1 or self.world == 1 or self._vec_dtype == dtype or self.peer_vec:
            return
        X, Y, _ = self.shape
        self._vec_dtype = dtype
        mk = lambda: torch.zeros((3, X, Y, self.vh), dtype=dtype, device=self.dev)
        self.vhalo_lo = mk() if self.rank > 0 else None                    # planes [z0 - vh, z0) of the lower neighbour
        self.vhalo_hi = mk() if self.rank < self.world - 1 else None       # planes [z1, z1 + vh) of the upper neighbour
        self.vsend_lo = mk() if self.rank > 0 else None                    # contiguous copies of my own faces, as they travel
        self.vsend_hi = mk() if self.rank < self.world - 1 else None

    def pack_vector_faces(self) -> None:
        """my first / last vh planes of the vector field as contiguous (3,X,Y,vh) blocks (one strided copy each)."""
        if self.vsend_lo is not None:
            self.vsend_lo.copy_(self.vec[:, :, :, :self.vh])
        if self.vsend_hi is not None:
            self.vsend_hi.copy_(self.vec[:, :, :, self.Zl - self.vh:])

    def exchange_vector_halos(self) -> None:
        """N > 1: swap vh planes of the vector field with both Z-neighbours (my low face -> the lower rank's high halo, my
        high face -> the upper rank's low halo).  NCCL send/recv for both transports: 2 x 3 x X x Y x vh elements per face."""
        if self.hops == 1 or self.world == 1 or self.peer_vec:
            return
        self.pack_vector_faces()
        # the same routing as the boundary runs (tests/test_sharded_gloo.py): my low face -> rank-1's high halo, ...
        self.comm.neighbour_exchange(self.vsend_lo, self.vsend_hi, self.vhalo_lo, self.vhalo_hi)

    def _s(self):
        return L.stream_ptr(self.dev)

    def _chain(self):
        """context of the labelling chain: the high-priority side stream when the gather is split."""
        if not self.split:
            return contextlib.nullcontext()
        from .pipeline import chain_stream
        return torch.cuda.stream(chain_stream(self.dev))

    def _label_local(self, phase: int) -> None:
        X, Y, Z = self.shape
        # the status word is sticky across passes (graph replays included): check_status() reads and clears it
        flags = phase | L.CCL_KEEP_STATUS | (L.CCL_WORKSPACE_CLEAN if self._clean and phase != L.CCL_PHASE_LABEL else 0)
        L.check(self.lib.skb_shard_label_local(self.mask.data_ptr(), L.dtype_code(self.mask), X, Y, Z, self.z_range[0], self.Zl,
                                               self.capacity, self.workspace.data_ptr(), self.workspace.numel(),
                                               self.meta[1:2].data_ptr(), flags, self._s()))

    def _stream_phase(self) -> None:
        X, Y, Z = self.shape
        L.check(self.lib.skb_assemble_stream(self.vec.data_ptr(), L.dtype_code(self.vec), X, Y, Z, self.z_range[0], self.Zl,
                                             self.workspace.data_ptr(), self.flags.data_ptr(), self.out.data_ptr(),
                                             L.dtype_code(self.out), self.stream_ctas, self._s()))

    # ---- phases ------------------------------------------------------------------------------------
    def phase_local(self) -> None:
        X, Y, Z = self.shape
        z0, z1 = self.z_range
        peer = self.transport == "peer"
        with torch.cuda.device(self.dev):
            # the halo words of the previous pass are cleared from its run lists (not a 33 MB memset per face and pass);
            # this has to happen before the next exchange overwrites the lists
            if peer:  # pass counter + both faces' halo words + the exchange buffer's counters: two launches
                L.check(self.lib.skb_shard_begin_pass(self.mailbox.ptr, *self._geom, Z, L.ptr(self.halo_lo), L.ptr(self.halo_hi),
                                                      self.exch.data_ptr(), self._s()))
            else:
                for hi, halo in ((0, self.halo_lo), (1, self.halo_hi)):
                    if halo is not None:
                        L.check(self.lib.skb_shard_clear_halo(Z, (self.recv_hi if hi else self.recv_lo).data_ptr(), self.cap_runs,
                                                              halo.data_ptr(), self._s()))
            if self.split:
                from .pipeline import chain_stream
                main = torch.cuda.current_stream(self.dev)
                self._label_local(L.CCL_PHASE_PACK)
                bits_ready = torch.cuda.Event()
                bits_ready.record(main)
                self._stream_phase()  # owns the current stream from here until the resolve
                chain_stream(self.dev).wait_event(bits_ready)
            else:
                self._label_local(0)
            self._clean = False
        with torch.cuda.device(self.dev), self._chain():
            if self.split:
                self._label_local(L.CCL_PHASE_LABEL)
            if peer:
                L.check(self.lib.skb_shard_emit_runs_peer(
                    self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, self.halo, self.mailbox.ptr,
                    self.peer_ptrs[self.rank - 1] if self.rank > 0 else 0,
                    self.peer_ptrs[self.rank + 1] if self.rank < self.world - 1 else 0,
                    *self._geom, self.meta[1:2].data_ptr(), self._s()))
                return
            if self.rank > 0:
                L.check(self.lib.skb_shard_emit_runs(self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, 0, self.halo,
                                                     self.send_lo.data_ptr(), self.cap_runs, self.meta[1:2].data_ptr(), self._s()))
            if self.rank < self.world - 1:
                L.check(self.lib.skb_shard_emit_runs(self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, 1, self.halo,
                                                     self.send_hi.data_ptr(), self.cap_runs, self.meta[1:2].data_ptr(), self._s()))

    def phase_ingest(self) -> None:
        """after the neighbour exchange: recv_lo / recv_hi hold the neighbours' boundary runs."""
        X, Y, Z = self.shape
        z0, z1 = self.z_range
        peer = self.transport == "peer"
        with torch.cuda.device(self.dev), self._chain():
            if not peer:  # zero the exchange buffer's counters and pack my roots (needs only the local phase) ...
                L.check(self.lib.skb_shard_boundary_pairs(self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, self.capacity,
                                                          0, self.exch.data_ptr(), self.cap_roots, self.cap_pairs,
                                                          self.meta[1:2].data_ptr(), self._s()))
            # ... then the neighbours' runs; the upper neighbour's ingest also appends the face pairs
            for hi, halo, recv in ((0, self.halo_lo, None if peer else self.recv_lo),
                                   (1, self.halo_hi, None if peer else self.recv_hi)):
                if halo is None:
                    continue
                exch = self.exch.data_ptr() if hi else 0
                if peer:
                    L.check(self.lib.skb_shard_ingest_runs_peer(self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, self.mailbox.ptr,
                                                                hi, *self._geom, halo.data_ptr(), exch,
                                                                self.meta[1:2].data_ptr(), self._s()))
                else:
                    L.check(self.lib.skb_shard_ingest_runs(self.workspace.data_ptr(), X, Y, Z, z0, self.Zl, recv.data_ptr(),
                                                           self.cap_runs, halo.data_ptr(), exch, self.cap_roots, self.cap_pairs,
                                                           self.meta[1:2].data_ptr(), self._s()))
            if peer:
                # roots straight from the slab's root list + the pairs the ingest appended -> every rank's mailbox
                L.check(self.lib.skb_shard_push(self.workspace.data_ptr(), X, Y, Z, self.capacity, self.exch.data_ptr(), self.mailbox.ptr,
                                                self._peer_arr, self.world, self.rank, self.cap_runs, self.cap_roots,
                                                self.cap_pairs, self.meta[1:2].data_ptr(), self._s()))

    def phase_merge(self) -> None:
        """after the all-gather: `gathered` holds every rank's roots and pairs -> global numbering on this rank."""
        X, Y, Z = self.shape
        with torch.cuda.device(self.dev), self._chain():
            if self.transport == "peer":
                L.check(self.lib.skb_shard_merge_peer(self.workspace.data_ptr(), X, Y, Z, self.capacity, self.mailbox.ptr,
                                                      self.world, self.rank, self.cap_runs, self.cap_roots, self.cap_pairs, 2,
                                                      self.meta[0:1].data_ptr(), self.meta[1:2].data_ptr(), self._s()))
            else:
                L.check(self.lib.skb_shard_merge(self.workspace.data_ptr(), X, Y, Z, self.capacity, self.gathered.data_ptr(),
                                                 self.world, self.rank, self.cap_roots, self.cap_pairs, 2,
                                                 self.meta[0:1].data_ptr(), self.meta[1:2].data_ptr(), self._s()))
            self._clean = True  # a completed merge leaves the root bitmap zeroed
            if self.split:
                self._chain_done = torch.cuda.Event()
                self._chain_done.record(torch.cuda.current_stream(self.dev))

    def gather(self, voxel_range: Optional[Tuple[int, int]] = None) -> Tensor:
        """the fused slab gather (or, split, the resolve pass) over the slab or a stretch [first, first+count) of its
        flat (X,Y,Zl) index.  A target beyond the planes the neighbours' runs describe sets STATUS_HALO_RANGE."""
        X, Y, Z = self.shape
        z0, _ = self.z_range
        with torch.cuda.device(self.dev):
            if self.split:
                torch.cuda.current_stream(self.dev).wait_event(self._chain_done)
                L.check(self.lib.skb_assemble_resolve(self.vec.data_ptr(), L.dtype_code(self.vec), X, Y, Z, z0, self.Zl,
                                                      L.f3(self.scale), self.workspace.data_ptr(), L.ptr(self.halo_lo),
                                                      L.ptr(self.halo_hi), self.flags.data_ptr(), self.out.data_ptr(),
                                                      L.dtype_code(self.out), self._s()))
                return self.out
            first, count = (0, X * Y * self.Zl) if voxel_range is None else (int(voxel_range[0]), int(voxel_range[1]))
            L.check(self.lib.skb_assemble_slab_ex(
                self.vec.data_ptr(), L.dtype_code(self.vec), X, Y, Z, z0, self.Zl, L.f3(self.scale), self.hops, self.decay,
                self._crop, self._overlap, self._vptr(self.vhalo_lo), self._vptr(self.vhalo_hi), self.vh, self.vdepth_lo, self.vdepth_hi,
                self.workspace.data_ptr(),
                L.ptr(self.halo_lo), L.ptr(self.halo_hi), self.halo, self.out.data_ptr(), L.dtype_code(self.out), first, count,
                self.meta[1:2].data_ptr(), self._s()))
        return self.out

    @staticmethod
    def _vptr(h) -> int:
        return 0 if h is None else (int(h) if isinstance(h, int) else h.data_ptr())

    @property
    def graphable(self) -> bool:
        """a pass has no host step or NCCL call inside: N = 1, or N > 1 with the neighbours' vector slabs mapped"""
        return self.transport == "peer" and (self.hops == 1 or self.peer_vec)

    def phase_merge_and_gather(self, timers=None) -> Tensor:
        self.phase_merge()
        if timers is not None and not self.split:
            timers[0].record()
        out = self.gather()
        if timers is not None and not self.split:
            timers[1].record()
        return out

    # ---- one pass over torch.distributed ----------------------------------------------------------
    def capture(self) -> bool:
        """Records one whole pass — ~25 kernel launches, the NCCL send/recv and the all-gather — into a CUDA
        graph; `step()` then replays it with a single launch.  At 8 GPUs a pass is ~1 ms of device time and
        issuing it call by call from Python costs about as much, so the host becomes the bottleneck.
        Needs a few eager passes first (NCCL communicators must exist).  Returns False (and stays eager)
        if the capture is refused."""
        self.graph = None
        try:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager(None)
            torch.cuda.synchronize(self.dev)
            self.graph = g
        except Exception as exc:  # pragma: no cover - depends on the NCCL / driver combination
            self.graph_error = repr(exc)
            self.graph = None
            torch.cuda.synchronize(self.dev)
        return self.graph is not None

    def step(self, timers=None, check: bool = True) -> Tensor:
        """one pass over the loaded slab.  check=True (default) reads the 8-byte status word back before returning the
        labels, so a peer that fell out of step (in-kernel waits are bounded, not infinite), an overflowed list or a
        gather target beyond the halo raises here instead of returning labels built from partial data.  A caller that
        times a loop of passes skips the read (check=False) and calls `check_status()` once after the loop."""
        if getattr(self, "graph", None) is not None and timers is None:
            self.graph.replay()
        else:
            self._step_eager(timers)
        if check:
            self.check_status()
        return self.out

    def check_status(self) -> int:
        """raises on any status bit of the last pass(es); returns the component count.  One small D2H read (synchronises)."""
        ncomp, status = (int(v) for v in self.meta.tolist())
        if status:
            self.meta[1:2].zero_()
        if status & L.STATUS_ROOT_OVERFLOW:
            raise L.SkootsB200Error("sharded CCL: a list capacity (roots / runs / pairs) overflowed")
        if status & L.STATUS_PEER_TIMEOUT:
            raise L.SkootsB200Error("sharded pass: a peer's flag did not arrive (a rank fell out of step or died)")
        if status & L.STATUS_HALO_RANGE:
            raise L.SkootsB200Error(f"sharded gather: a vector points more than {self.halo} planes beyond this rank's slab "
                                    "(|v_z * scale_z| exceeds the halo the neighbours sent); construct the assembler with a larger halo")
        return ncomp

    def _step_eager(self, timers=None) -> Tensor:
        self.exchange_vector_halos()  # N > 1 only; first in the pass on every rank, before any kernel waits on a peer's flag
        self.phase_local()
        if self.transport != "peer":
            with self._chain():
                self.comm.neighbour_exchange(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi)
        self.phase_ingest()
        if self.transport != "peer":
            with self._chain():
                self.comm.all_gather(self.gathered, self.exch)
        return self.phase_merge_and_gather(timers)

    def time_dominant(self, steps: int = 5) -> Tuple[str, float]:
        """(kernel name, mean ms per launch) of the pass's dominant kernel timed ALONE with CUDA events on this
        rank's slab: the stream phase when the gather is split, else the fused slab gather (run after a pass,
        so the labels it reads exist)."""
        X, Y, Z = self.shape
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        with torch.cuda.device(self.dev):
            for _ in range(steps + 1):
                a.record()
                if self.split:
                    self._stream_phase()
                else:
                    self.gather()
                b.record()
                torch.cuda.synchronize(self.dev)
                if _:
                    total += a.elapsed_time(b)
            if self.split:  # the stream phase left the flagged groups unwritten: finish the pass it started
                L.check(self.lib.skb_assemble_resolve(self.vec.data_ptr(), L.dtype_code(self.vec), X, Y, Z, self.z_range[0], self.Zl,
                                                      L.f3(self.scale), self.workspace.data_ptr(), L.ptr(self.halo_lo),
                                                      L.ptr(self.halo_hi), self.flags.data_ptr(), self.out.data_ptr(),
                                                      L.dtype_code(self.out), self._s()))
                torch.cuda.synchronize(self.dev)
        return ("assemble_stream_kernel" if self.split else "assemble_kernel"), total / steps

    def profile_phases(self, steps: int = 5) -> dict:
        """mean device time (ms) of each phase of a pass on this rank (CUDA events on the current stream).
        Peer transport: the waits for the other ranks sit inside `ingest_pairs` and `merge`."""
        peer = self.transport == "peer"
        names = ["local", "exchange_runs", "ingest_pairs", "all_gather", "merge", "gather"]
        acc = dict.fromkeys(names, 0.0)
        saved_split, self.split = self.split, False  # phases back to back on one stream, fused gather
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            mid = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            self.phase_local(); ev[1].record()
            if not peer:
                self.comm.neighbour_exchange(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi)
            ev[2].record()
            self.phase_ingest(); ev[3].record()
            if not peer:
                self.comm.all_gather(self.gathered, self.exch)
            ev[4].record()
            self.phase_merge_and_gather(mid); ev[6].record()
            torch.cuda.synchronize(self.dev)
            spans = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
                     ev[3].elapsed_time(ev[4]), ev[4].elapsed_time(mid[0]), mid[0].elapsed_time(mid[1])]
            for n, v in zip(names, spans):
                acc[n] += v / steps
        self.split = saved_split
        if not peer:
            acc["runs_sent"] = [int(self.send_lo[0].item()), int(self.send_hi[0].item())]
        acc["roots_pairs"] = [int(self.exch[0].item()), int(self.exch[1].item())]
        return acc

    def check(self) -> Tuple[int, int]:
        """(n_components, labelled voxels over all ranks); raises on a capacity overflow."""
        ncomp = self.check_status()
        labelled = (self.out > 0).sum().to(torch.int64)
        if self.comm is not None and self.world > 1:
            self.comm.dist.all_reduce(labelled, group=self.comm.group)
        return ncomp, int(labelled.item())

    # ---- host buffers -----------------------------------------------------------------------------------
    def _host_plan(self, n_slabs: int):
        """X-ranges of the pipelined host pass; a range starts on a multiple of 256 slab voxels (skb_assemble_slab_ex)."""
        X, Y, _ = self.shape
        plane = Y * self.Zl
        bounds = sorted({X * i // n_slabs for i in range(n_slabs + 1)})
        bounds = [b for b in bounds if (b * plane) % 256 == 0 or b == X]
        if bounds[0] != 0:
            bounds = [0] + bounds
        return list(zip(bounds[:-1], bounds[1:]))

    def run_host(self, mask_host: Tensor, vec_host: Tensor, out_host: Tensor, n_slabs: int = 8, check: bool = True) -> Tensor:
        """One pass with this rank's slab in HOST memory (pinned for full speed): mask (X,Y,Zl) u8, vectors (3,X,Y,Zl),
        labels written to out_host (X,Y,Zl) int32 | int16 — the sharded form of `pipeline.HostAssembler`.

        Pipelined on three streams: the mask goes up first and the whole labelling chain (local labelling, both
        exchanges with the other ranks, merge) runs while the vector field follows X-range by X-range; each range is
        gathered as soon as it has landed and its labels travel back while the next range is still arriving, so both
        PCIe directions and the kernels overlap.  The pass ends with a device synchronisation and the status check."""
        X, Y, Z = self.shape
        if self.split:
            raise L.SkootsB200Error("run_host pipelines the fused slab gather; construct the assembler without split=True")
        if self.mask is None or self.mask.dtype != torch.uint8 or tuple(self.mask.shape) != (X, Y, self.Zl):
            self.mask = torch.empty((X, Y, self.Zl), dtype=torch.uint8, device=self.dev)
        if self.peer_vec:
            self._peer_vector_slab(vec_host.dtype)
            torch.cuda.synchronize(self.dev)
            self.comm.barrier()  # nobody's walks of the previous pass still read the slab this upload overwrites
        elif self.vec is None or self.vec.dtype != vec_host.dtype:
            self.vec = torch.empty((3, X, Y, self.Zl), dtype=vec_host.dtype, device=self.dev)
        self._alloc_vector_halos(self.vec.dtype)
        if out_host.dtype != self.out.dtype:
            self.graph = None  # a captured pass writes the old buffer
            self.out = torch.empty((X, Y, self.Zl), dtype=out_host.dtype, device=self.dev)
        if getattr(self, "_up", None) is None:
            self._up, self._down = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        mask_host, vec_host, out_host = mask_host.reshape(X, Y, self.Zl), vec_host.reshape(3, X, Y, self.Zl), out_host.reshape(X, Y, self.Zl)
        plane = Y * self.Zl
        main = torch.cuda.current_stream(self.dev)
        self._up.wait_stream(main)
        self._down.wait_stream(main)
        ranges = self._host_plan(n_slabs) if self.hops == 1 else [(0, X)]  # N > 1: a walk may read any X-range of its crop
        landed = []
        with torch.cuda.stream(self._up):
            self.mask.copy_(mask_host, non_blocking=True)
            mask_ready = torch.cuda.Event()
            mask_ready.record(self._up)
            for x0, x1 in ranges:
                for c in range(3):  # one contiguous block per channel and range: a plain DMA each
                    self.vec[c, x0:x1].copy_(vec_host[c, x0:x1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._up)
                landed.append(ev)
        main.wait_event(mask_ready)
        self.phase_local()
        if self.transport != "peer":
            self.comm.neighbour_exchange(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi)
        self.phase_ingest()
        if self.transport != "peer":
            self.comm.all_gather(self.gathered, self.exch)
        self.phase_merge()
        if self.hops > 1 and self.peer_vec:   # the walks read the neighbours' slabs in place: every upload must have landed
            landed[-1].synchronize()
            self.comm.barrier()
        elif self.hops > 1:  # the walks read the neighbours' vectors: swap the faces once the field has landed (every rank
            main.wait_event(landed[-1])  # issues it at this point of its pass, after the chain's in-kernel waits)
            self.exchange_vector_halos()
        for (x0, x1), ev in zip(ranges, landed):
            main.wait_event(ev)
            self.gather((x0 * plane, (x1 - x0) * plane))
            done = torch.cuda.Event()
            done.record(main)
            self._down.wait_event(done)
            with torch.cuda.stream(self._down):
                out_host[x0:x1].copy_(self.out[x0:x1], non_blocking=True)
        torch.cuda.synchronize(self.dev)
        if check:
            self.check_status()
        return out_host

    def e2e(self, steps: int, mask_host: Optional[Tensor] = None, vec_host: Optional[Tensor] = None,
            out_host: Optional[Tensor] = None, out_dtype=None, n_slabs: int = 8) -> dict:
        """times `run_host` over `steps` passes (wall clock between barriers, max over ranks) with this rank's slab in
        pinned HOST memory; returns the e2e record of bench.py plus this rank's PCIe rates."""
        import time
        X, Y, Z = self.shape
        with L.numa_local(self.dev) as numa:  # this rank's pinned buffers on its GPU's NUMA node
            if mask_host is None:
                mask_host = self.mask.cpu().pin_memory()
                vec_host = self.vec.cpu().pin_memory()
            if out_host is None:
                out_host = torch.empty((X, Y, self.Zl), dtype=out_dtype or self.out.dtype).pin_memory()
        self.run_host(mask_host, vec_host, out_host, n_slabs)
        self.comm.barrier()
        torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.run_host(mask_host, vec_host, out_host, n_slabs, check=False)
        mine = (time.perf_counter() - t0) / steps
        self.comm.barrier()
        wall = (time.perf_counter() - t0) / steps
        self.check_status()
        h2d = mask_host.numel() * mask_host.element_size() + vec_host.numel() * vec_host.element_size()
        d2h = out_host.numel() * out_host.element_size()
        dist = self.comm.dist
        t = torch.tensor([wall, mine], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.comm.group)
        dt = float(t[0].item())
        rates = torch.zeros(self.world, dtype=torch.float64, device=self.dev)
        rates[self.rank] = (h2d + d2h) / mine / 1e9
        dist.all_reduce(rates, group=self.comm.group)
        self.host_out = out_host
        return {"value": X * Y * Z / dt, "unit": "voxels/s", "h2d_bytes_per_step": h2d * self.world,
                "d2h_bytes_per_step": d2h * self.world, "ms_per_step": dt * 1e3, "steps": steps,
                "out_dtype": str(out_host.dtype).replace("torch.", ""),
                "pcie_GBps_per_rank": [round(float(v), 1) for v in rates.tolist()], "numa_rank0": numa,
                "pipeline": f"mask up -> labelling chain + exchanges while {len(self._host_plan(n_slabs))} X-ranges of the vectors "
                            "follow; each range gathered when landed, labels travel back meanwhile (3 streams)",
                "api": "skoots_b200.sharded.ShardedAssembler.run_host"}


class LocalGroup:
    """All ranks of a sharded pass inside ONE process on ONE GPU.  For tests (and for a box with fewer GPUs
    than ranks).  transport "peer": the ranks' mailboxes live on the same device, so "peer" pointers are
    ordinary pointers, and the phases are issued rank by rank so that every flag has been released before
    a kernel waits on it.  transport "nccl": the collectives become plain copies."""

    class _PeerTag:
        transport = "peer"

    def __init__(self, shape, world: int, device, scale=(60, 60, 12), transport: str = "peer", **kw):
        comm = self._PeerTag() if transport == "peer" else None
        self.transport = transport
        self.ranks = [ShardedAssembler(shape, world, r, device, scale=scale, comm=comm, **kw) for r in range(world)]
        self.world = world
        if transport == "peer":
            self.mailboxes = [Mailbox(r.lib, r.mailbox_bytes, device) for r in self.ranks]
            ptrs = [m.ptr for m in self.mailboxes]
            for r, m in zip(self.ranks, self.mailboxes):
                r.attach(m, ptrs)

    def load_volume(self, mask: Tensor, vec: Tensor) -> None:
        for r in self.ranks:
            z0, z1 = r.z_range
            r.load(mask[:, :, z0:z1].contiguous(), vec[:, :, :, z0:z1].contiguous())
        if self.ranks[0].peer_vec:  # same process: the neighbours' slabs are ordinary device pointers
            for i, r in enumerate(self.ranks):
                lo, hi = (self.ranks[i - 1] if i > 0 else None), (self.ranks[i + 1] if i < self.world - 1 else None)
                r.attach_vector_peers(lo.vec.data_ptr() if lo else 0, lo.Zl if lo else 0, hi.vec.data_ptr() if hi else 0, hi.Zl if hi else 0)

    def step(self) -> Tensor:
        if self.ranks[0].hops > 1 and self.world > 1 and not self.ranks[0].peer_vec:  # stand-in for exchange_vector_halos()
            for r in self.ranks:
                r.pack_vector_faces()
            for i, r in enumerate(self.ranks):
                if i > 0:
                    r.vhalo_lo.copy_(self.ranks[i - 1].vsend_hi)
                if i < self.world - 1:
                    r.vhalo_hi.copy_(self.ranks[i + 1].vsend_lo)
        for r in self.ranks:
            r.phase_local()
        if self.transport != "peer":
            with self.ranks[0]._chain():  # the stand-ins for the collectives go where the chain runs
                for i, r in enumerate(self.ranks):
                    if i > 0:
                        r.recv_lo.copy_(self.ranks[i - 1].send_hi)
                    if i < self.world - 1:
                        r.recv_hi.copy_(self.ranks[i + 1].send_lo)
        for r in self.ranks:
            r.phase_ingest()
        if self.transport != "peer":
            with self.ranks[0]._chain():
                allg = torch.cat([r.exch for r in self.ranks])
                for r in self.ranks:
                    r.gathered.copy_(allg)
        outs = [r.phase_merge_and_gather() for r in self.ranks]
        for r in self.ranks:
            r.check_status()
        return torch.cat(outs, dim=2)
