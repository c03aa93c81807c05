"""GPU, >= 2 devices: the REAL multi-process peer path — one process per GPU, CUDA-IPC mailboxes, kernels storing
into the other GPU's memory over NVLink and spinning on flags — must be bit-identical to the unsharded pass, eagerly,
as a replayed CUDA graph, through the NCCL transport and through the pipelined host-buffer pass.  (tests/test_gpu_sharded.py
covers the same phases with all ranks emulated in one process; this file is skipped on a single-GPU box and is run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multiproc.py` — log under profiles/.)"""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _volume(shape, dev):
    from skoots_b200.synthetic import make_tube_volume
    tv = make_tube_volume(shape, 160, seed=3, device=dev)
    tv.skeleton[40:43, 30:33, 20:shape[2] - 16] = 1      # one object through every slab face
    tv.vectors[2, 36:47, 26:37, :] = 0.5                 # vectors that cross the faces
    return tv.skeleton, tv.vectors


def _worker(rank: int, world: int, port: int, transport: str, shape):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.sharded import PeerComm, ShardedAssembler, TorchDistComm
    scale = (60, 60, 12)
    mask, vec = _volume(shape, dev)
    want_full = assemble_instances(mask, vec, torch.tensor(scale), N=1)
    run = ShardedAssembler(shape, world, rank, dev, scale=scale, comm=PeerComm() if transport == "peer" else TorchDistComm())
    z0, z1 = run.z_range
    want = want_full[:, :, z0:z1].contiguous()
    run.load(mask[:, :, z0:z1].contiguous(), vec[:, :, :, z0:z1].contiguous())
    dist.barrier()
    for _ in range(3):                                   # eager passes (both copies of every receive buffer)
        assert torch.equal(run.step(), want), f"rank {rank}: eager sharded pass differs from the unsharded one"
    ncomp, labelled = run.check()
    assert ncomp == int(want_full.max()) - 2 and labelled == int((want_full > 0).sum())
    if transport == "peer":
        ok = torch.tensor([1 if run.capture() else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assert int(ok.item()) == 1, f"rank {rank}: the peer pass did not capture into a CUDA graph"
        for _ in range(4):
            run.out.zero_()
            assert torch.equal(run.step(), want), f"rank {rank}: replayed graph differs"
    # host buffers: pipelined upload / chain / gather / download, int32 and the reference's int16
    mask_h, vec_h = run.mask.cpu().pin_memory(), run.vec.cpu().pin_memory()
    for dt in (torch.int32, torch.int16):
        out_h = torch.empty(want.shape, dtype=dt).pin_memory()
        for _ in range(2):
            out_h.fill_(-1)
            run.run_host(mask_h, vec_h, out_h, n_slabs=4)
            assert torch.equal(out_h, want.cpu().to(dt)), f"rank {rank}: host pass ({dt}) differs"
    dist.barrier()
    # eval()'s configuration Z-sharded: crop grid, N = 6 hops.  NCCL transport: packed copies of the faces' vector planes
    # travel by send/recv every pass; peer transport: the neighbours' vector slabs are mapped through CUDA IPC and the
    # hops that cross a face read the other GPU's memory directly
    crop, ov = (48, 40, 50), (4, 4, 5)
    want_eval = assemble_instances(mask, vec, torch.tensor(scale), N=6, crop=crop, overlap=ov, out_dtype=torch.int16)[:, :, z0:z1].contiguous()
    comm2 = PeerComm() if transport == "peer" else run.comm
    ev = ShardedAssembler(shape, world, rank, dev, scale=scale, hops=6, crop=crop, overlap=ov, comm=comm2, out_dtype=torch.int16)
    assert ev.peer_vec == (transport == "peer")
    ev.load(run.mask, run.vec)
    for _ in range(2):
        assert torch.equal(ev.step(), want_eval), f"rank {rank}: sharded eval-mode pass differs from the unsharded one"
    if ev.graphable:
        ok = torch.tensor([1 if ev.capture() else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assert int(ok.item()) == 1
        for _ in range(2):
            ev.out.zero_()
            assert torch.equal(ev.step(), want_eval), f"rank {rank}: replayed eval-mode graph differs"
    out_h = torch.empty(want_eval.shape, dtype=torch.int16).pin_memory()
    for _ in range(2):
        out_h.fill_(-1)
        ev.run_host(mask_h, vec_h, out_h)
        assert torch.equal(out_h, want_eval.cpu()), f"rank {rank}: eval-mode host pass differs"
    dist.barrier()
    if comm2 is not run.comm and hasattr(comm2, "close"):
        ev.graph = None
        comm2.close()
    if hasattr(run.comm, "close"):
        run.comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_processes_bit_identical_to_unsharded(transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    shape = (96, 80, 64 * world * 2)
    mp.spawn(_worker, args=(world, _free_port(), transport, shape), nprocs=world, join=True)
