"""CPU, world_size 2 over gloo: the host-side plumbing of the sharded pass — slab partition, exchange
buffer layout, and that TorchDistComm routes boundary runs to the right neighbour and all-gathers
payloads in rank order.  (The kernels themselves need a GPU: tests/test_gpu_sharded.py.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skoots_b200.sharded import TorchDistComm, exchange_layout, slab_bounds


def test_slab_bounds():
    assert slab_bounds(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]
    assert slab_bounds(512, 2) == [(0, 256), (256, 512)]
    assert slab_bounds(192, 2) == [(0, 64), (64, 192)]
    with pytest.raises(ValueError):
        slab_bounds(500, 2)
    with pytest.raises(ValueError):
        slab_bounds(128, 4)
    assert exchange_layout(8, 4) == 2 + 8 + 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchDistComm()
        n = 12
        send_lo = torch.full((n,), 10 * rank + 1, dtype=torch.int32)
        send_hi = torch.full((n,), 10 * rank + 2, dtype=torch.int32)
        recv_lo = torch.full((n,), -1, dtype=torch.int32)
        recv_hi = torch.full((n,), -1, dtype=torch.int32)
        comm.neighbour_exchange(send_lo, send_hi, recv_lo, recv_hi)
        payload = torch.arange(5, dtype=torch.int32) + 100 * rank
        gathered = torch.empty(5 * world, dtype=torch.int32)
        comm.all_gather(gathered, payload)
        comm.barrier()
        results[rank] = (recv_lo.tolist(), recv_hi.tolist(), gathered.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_comm_routing_gloo(world):
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        for r in range(world):
            lo, hi, gathered = results[r]
            assert lo == ([10 * (r - 1) + 2] * 12 if r > 0 else [-1] * 12)        # lower neighbour's send_hi
            assert hi == ([10 * (r + 1) + 1] * 12 if r < world - 1 else [-1] * 12)  # upper neighbour's send_lo
            assert gathered == [v + 100 * q for q in range(world) for v in range(5)]


@pytest.mark.parametrize("shape,Zl,n", [((2048, 2048, 512), 64, 8), ((7, 5, 128), 64, 8), ((1, 3, 64), 64, 8), ((33, 17, 192), 128, 4),
                                        ((100, 9, 64), 64, 16), ((5, 1, 64), 64, 3)])
def test_host_pipeline_ranges_cover_the_slab_and_start_on_gather_boundaries(shape, Zl, n):
    """run_host's X-ranges: contiguous, covering [0, X), each starting on a multiple of 256 slab voxels (what
    skb_assemble_slab_ex accepts as the start of a range), whatever the shape."""
    import types
    from skoots_b200.sharded import ShardedAssembler
    ranges = ShardedAssembler._host_plan(types.SimpleNamespace(shape=shape, Zl=Zl), n)
    X, Y, _ = shape
    assert ranges and ranges[0][0] == 0 and ranges[-1][1] == X
    assert all(a1 == b0 for (_, a1), (b0, _) in zip(ranges[:-1], ranges[1:]))
    assert all(x1 > x0 and (x0 * Y * Zl) % 256 == 0 for x0, x1 in ranges)
    assert len(ranges) <= n
