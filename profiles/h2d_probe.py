#!/usr/bin/env python
"""How fast can N ranks pull host memory at the same time?  (The end-to-end pass at 8 GPUs is bound by exactly this:
profiles/README.md.)  Every rank copies the same amount host -> device (and device -> host) while all others do the same;
variants: torch's pinned memory (cudaHostAlloc default), write-combined pinned memory, one or two copy streams.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/h2d_probe.py
"""
import ctypes
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 1 << 30
rt = ctypes.CDLL("libcudart.so.12")


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, f"cudaHostAlloc rc={rc}"
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p


def bench(host, dev_buf, n_streams, direction):
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    chunk = NBYTES // n_streams
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_event(a)
    for rep in range(4):
        for k, st in enumerate(streams):
            with torch.cuda.stream(st):
                sl = slice(k * chunk, (k + 1) * chunk)
                if direction == "h2d":
                    dev_buf[sl].copy_(host[sl], non_blocking=True)
                elif direction == "d2h":
                    host[sl].copy_(dev_buf[sl], non_blocking=True)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    return 4 * NBYTES / (a.elapsed_time(b) * 1e-3) / 1e9


dev_buf = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
out = {}
for name, flags in (("pinned", 0), ("write_combined", 4)):   # cudaHostAllocWriteCombined = 0x04
    host, keep = host_alloc(NBYTES, flags)
    host.fill_(1) if name == "pinned" else host.copy_(torch.ones(NBYTES, dtype=torch.uint8))
    for n_streams in (1, 2):
        for direction in ("h2d", "d2h"):
            if name == "write_combined" and direction == "d2h":
                continue
            bench(host, dev_buf, n_streams, direction)
            out[f"{name}_{direction}_{n_streams}s"] = round(bench(host, dev_buf, n_streams, direction), 1)
    rt.cudaFreeHost(keep)
vals = torch.tensor(list(out.values()), dtype=torch.float64, device=dev)
if world > 1:
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
else:
    gathered = [vals]
if rank == 0:
    res = {k: [round(float(g[i]), 1) for g in gathered] for i, k in enumerate(out)}
    res["aggregate_GBps"] = {k: round(sum(v), 1) for k, v in res.items()}
    print(json.dumps({"world": world, "GBps_per_rank": res}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
