#!/usr/bin/env python
"""bench.py — post-processing throughput of the skeleton-embedding instance-assembly path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the hot path (connected-component labelling of the u8 skeleton mask +
fused vector->embedding->label gather) over the synthetic analytic-tube volume named in
`config.workload`.  `value` is voxels/s with inputs resident in HBM (CUDA events, max over
ranks); `e2e` is the same pass through the public host-buffer API with the host<->device copies
inside the timed region.  See DESIGN.md §Measurement for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SCALE = (60, 60, 12)           # SKOOTS.VECTOR_SCALING default, skoots/config.py:144
ALGO_BYTES_PATH = 11.0         # u8 mask + 3 x fp16 vectors in, int32 label out
ALGO_BYTES_GATHER = 10.0       # dominant kernel (fused gather, or its stream phase): 6 B vectors in + 4 B labels out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default=os.environ.get("SKB_BENCH_SHAPE", "2048,2048,512"))
    ap.add_argument("--tubes", type=int, default=0, help="0 = 16384 scaled by volume")
    ap.add_argument("--hops", type=int, default=1, help="N of vector_to_embedding (eval() uses 10)")
    ap.add_argument("--mode", default="whole", choices=["whole", "eval"],
                    help="whole = lib functions on the whole volume (headline); eval = eval()'s 500/500/50 crop grid, "
                         "50/50/5 overlap, int16 labels (use with --hops 10 to replay skoots/lib/eval.py:245-284)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def n_tubes_for(shape, requested):
    if requested:
        return requested
    vox = shape[0] * shape[1] * shape[2]
    return max(8, int(round(16384 * vox / (2048 * 2048 * 512))))


def measured_traffic(kernel: str, voxels_per_launch: float):
    """dram read+write bytes of one launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/r01_traffic.json; captured at 2048x2048x512 on one GPU, scaled per voxel) — None if the capture is
    of another kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            rec = json.load(fh)
        if rec["kernel"].split("<")[0] != kernel:
            return None
        return rec["dram_bytes_per_voxel"] * voxels_per_launch
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock + throttle reasons through NVML every ~2 ms on a background thread."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        import threading
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover
            self.err = f"nvml unavailable: {exc}"
            self._nv = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        if self._nv is None:
            return
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception as exc:  # pragma: no cover
                self.err = str(exc)
                return
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        self._thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["note"] = self.err
        return out


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample_shape(budget_s: float):
    """largest sample of the workload whose oracle pass fits the budget (~12 Mvox/s at N=1)."""
    for shape in ((1024, 1024, 128), (768, 768, 128), (512, 512, 128), (384, 384, 128), (256, 256, 128), (128, 128, 64)):
        if shape[0] * shape[1] * shape[2] / 12e6 <= budget_s:
            return shape
    return (128, 128, 32)


EVAL_CROP, EVAL_OVERLAP = (500, 500, 50), (50, 50, 5)  # skoots/lib/eval.py:248-249


def mode_kwargs(mode):
    return dict(crop=EVAL_CROP, overlap=EVAL_OVERLAP) if mode == "eval" else dict(crop=None, overlap=(0, 0, 0))


def cpu_pass(mask, vec, hops, mode="whole"):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import skoots_oracle as orc
    t0 = time.perf_counter()
    kw = mode_kwargs(mode)
    out = orc.postprocess(mask, vec, torch.tensor(SCALE), N=hops, crop=kw["crop"], overlap=kw["overlap"],
                          out_dtype=torch.int16 if mode == "eval" else torch.int32)
    return time.perf_counter() - t0, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from skoots_b200.synthetic import make_tube_volume
    torch.set_num_threads(os.cpu_count() or 1)
    total_steps = max(1, args.steps + args.warmup)
    shape = cpu_sample_shape(150.0 / total_steps / (1.0 if args.hops == 1 else 12.0 * args.hops))
    full = tuple(int(v) for v in args.shape.split(","))
    tv = make_tube_volume(shape, n_tubes_for(shape, 0), seed=0, want_mask=False, want_skeleton_dict=False)
    for _ in range(args.warmup):
        cpu_pass(tv.skeleton, tv.vectors, args.hops, args.mode)
    times = [cpu_pass(tv.skeleton, tv.vectors, args.hops, args.mode)[0] for _ in range(args.steps)]
    vox = shape[0] * shape[1] * shape[2]
    value = vox * len(times) / sum(times)
    sample = f"{shape[0]}x{shape[1]}x{shape[2]} sub-volume of the workload, same tube density, whole pass per step"
    line = {
        "impl": "reference", "metric": "post-proc voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(full, args.hops, args.mode), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(shape, hops, mode="whole"):
    how = "whole-volume" if mode == "whole" else "eval() crop grid 500/500/50 ov 50/50/5,"
    return (f"synthetic analytic tubes {shape[0]}x{shape[1]}x{shape[2]} (BASELINE.json configs[2]), {how} "
            f"flood fill + vector_to_embedding(N={hops}) + index_skeleton_by_embed")


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import skoots_b200._lib as L_
    from skoots_b200.lib.flood_fill import label_components
    from skoots_b200.pipeline import HostAssembler, gather_instances
    from skoots_b200.synthetic import make_tube_volume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(int(v) for v in args.shape.split(","))
    X, Y, Z = shape
    V = X * Y * Z
    scale = torch.tensor(SCALE)
    hbm_peak, peak_src = peaks()

    if world > 1:
        from skoots_b200.sharded import PeerComm, ShardedAssembler, TorchDistComm
        import skoots_b200._lib as L_err
        transport = os.environ.get("SKB_TRANSPORT", "peer")
        make = lambda comm: ShardedAssembler(shape, world, rank, dev, scale=SCALE, hops=args.hops, comm=comm,
                                             split=bool(os.environ.get("SKB_SPLIT")))
        try:
            runner = make(PeerComm() if transport == "peer" else TorchDistComm())
        except L_err.SkootsB200Error as exc:  # collective failure (every rank raises): CUDA IPC is not usable on this box
            if transport != "peer":
                raise
            if rank == 0:
                print(f"bench.py: {exc}; falling back to the NCCL transport", file=sys.stderr, flush=True)
            runner = make(TorchDistComm())
        z0, z1 = runner.z_range
        tv = make_tube_volume(shape, n_tubes_for(shape, args.tubes), seed=0, device=dev, z_range=(z0, z1),
                              want_mask=False, want_skeleton_dict=False)
        runner.load(tv.skeleton, tv.vectors)
        del tv

        def step(timers=None):
            return runner.step(timers)
        launches_per_step = runner.launches_per_step
    else:
        tv = make_tube_volume(shape, n_tubes_for(shape, args.tubes), seed=0, device=dev, want_mask=False,
                              want_skeleton_dict=False)
        mask, vec = tv.skeleton, tv.vectors
        kw = mode_kwargs(args.mode)
        out = torch.empty(shape, dtype=torch.int16 if args.mode == "eval" else torch.int32, device=dev)
        state = {"ws": None, "sparse": None}
        from skoots_b200.lib.flood_fill import new_sparse
        from skoots_b200.pipeline import assemble_split, split_eligible
        split = bool(os.environ.get("SKB_SPLIT")) and split_eligible(shape, vec, args.hops, kw["crop"], kw["overlap"])
        if split:
            state["sparse"] = new_sparse(shape, dev)
            state["ws"] = state["sparse"].workspace
            group_flags = torch.empty(V // 256, dtype=torch.int32, device=dev)

        def step(timers=None):
            if split:  # pack | stream phase next to the labelling chain on a second stream | resolve
                return assemble_split(mask, vec, SCALE, state["sparse"], out, group_flags=group_flags, timers=timers)
            sp = label_components(mask, label_base=2, workspace=state["ws"], check=False)
            state["ws"], state["sparse"] = sp.workspace, sp
            if timers is not None:
                timers[0].record()
            gather_instances(vec, scale, sp, N=args.hops, out=out, **kw)
            if timers is not None:
                timers[1].record()
            return out
        # init, pack, tile, boundary, flatten, scan x2, rank, clear, publish, gather (split: stream + resolve) (+memsets)
        launches_per_step = 12 if split else 11

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()  # every rank's slab is loaded before anyone's kernels start waiting on a peer's flags
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    phases = None
    if world == 1:
        state["sparse"].check()
        n_components = state["sparse"].num_components
        labelled = int((out > 0).sum().item())
    else:
        n_components, labelled = runner.check()
        if os.environ.get("SKB_PHASE_TIMING"):
            phases = runner.profile_phases()

    sampler = ClockSampler(local) if rank == 0 else None
    timers = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graphed = False
    if world > 1 and not os.environ.get("SKB_NO_GRAPH"):
        # all ranks must agree: a rank replaying a graph and a rank issuing eagerly would still match
        # collectives, but keep the measurement uniform
        ok = torch.tensor([1 if runner.capture() else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        graphed = bool(ok.item())
        if not graphed:
            runner.graph = None
        else:
            for _ in range(3):
                step()
    barrier()
    t_begin.record()
    for k in range(args.steps):
        step(None if graphed else timers[k])
    t_end.record()
    barrier()
    elapsed_ms = t_begin.elapsed_time(t_end)
    # the dominant kernel ALONE (CUDA events around single launches on the same buffers, right after the timed
    # region): inside a pass it runs next to the labelling chain / inside a graph, where events cannot bracket it
    if world > 1:
        dominant, gather_ms = runner.time_dominant(max(3, min(args.steps, 10)))
    elif split:
        dominant, gather_ms = "assemble_stream_kernel", 0.0
        lib, sp = L_.load(), state["sparse"]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        for k in range(reps + 1):
            a.record()
            L_.check(lib.skb_assemble_stream(vec.data_ptr(), L_.dtype_code(vec), X, Y, Z, 0, Z, sp.workspace.data_ptr(),
                                             group_flags.data_ptr(), out.data_ptr(), L_.dtype_code(out), 0, L_.stream_ptr(dev)))
            b.record()
            torch.cuda.synchronize(dev)
            gather_ms += a.elapsed_time(b) / reps if k else 0.0
        step()  # leave `out` complete again
        torch.cuda.synchronize(dev)
    else:
        dominant, gather_ms = "assemble_kernel", sum(a.elapsed_time(b) for a, b in timers) / args.steps
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([elapsed_ms, gather_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, gather_ms = t.tolist()
    ms_per_step = elapsed_ms / args.steps
    value = V / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the public API ------------------------------------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            host_mask = torch.empty(shape, dtype=torch.uint8).pin_memory()
            host_vec = torch.empty((3,) + shape, dtype=torch.float16).pin_memory()
            host_out = torch.empty(shape, dtype=out.dtype).pin_memory()
            host_mask.copy_(mask)
            host_vec.copy_(vec)
            del out
            state["ws"] = state["sparse"] = None
            del mask, vec
            torch.cuda.empty_cache()
            runner_h = HostAssembler(shape, dev, out_dtype=host_out.dtype)
            e2e_steps = max(1, min(args.steps, 5))
            runner_h(host_mask, host_vec, scale, host_out, N=args.hops, **kw)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                runner_h(host_mask, host_vec, scale, host_out, N=args.hops, **kw)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / e2e_steps
            e2e = {"value": V / dt, "unit": "voxels/s", "h2d_bytes_per_step": host_mask.numel() + host_vec.numel() * 2,
                   "d2h_bytes_per_step": host_out.numel() * host_out.element_size(), "ms_per_step": dt * 1e3, "steps": e2e_steps,
                   "api": "skoots_b200.pipeline.HostAssembler"}
            assert int((host_out > 0).sum().item()) == labelled, "e2e result differs from the device-resident run"
        else:
            e2e = runner.e2e(max(1, min(args.steps, 5)))

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample + parity of that sample ----------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sshape = cpu_sample_shape(20.0 / (1.0 if args.hops == 1 else 12.0 * args.hops))
        sshape = tuple(min(a, b) for a, b in zip(sshape, shape))
        torch.set_num_threads(os.cpu_count() or 1)
        if e2e is not None:
            smask = host_mask[:sshape[0], :sshape[1], :sshape[2]].contiguous()
            svec = host_vec[:, :sshape[0], :sshape[1], :sshape[2]].contiguous()
        else:
            smask = mask[:sshape[0], :sshape[1], :sshape[2]].contiguous().cpu()
            svec = vec[:, :sshape[0], :sshape[1], :sshape[2]].contiguous().cpu()
        cpu_pass(smask[:64, :64, :32].contiguous(), svec[:, :64, :64, :32].contiguous(), args.hops, args.mode)  # warm the thread pool
        dt, want = cpu_pass(smask, svec, args.hops, args.mode)
        from skoots_b200.pipeline import assemble_instances
        got = assemble_instances(smask.to(dev), svec.to(dev), scale, N=args.hops, out_dtype=want.dtype, **kw).cpu()
        parity = bool(torch.equal(got, want))
        svox = sshape[0] * sshape[1] * sshape[2]
        cpu = {"value": svox / dt, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"first {sshape[0]}x{sshape[1]}x{sshape[2]} voxels of the workload volume, one whole pass "
                         f"({dt:.1f} s); GPU result on the same sample bit-exact: {parity}"}
        assert parity, "GPU instance mask differs from the oracle on the CPU-baseline sample"

    if rank == 0:
        gather_bytes = ALGO_BYTES_GATHER * V / world
        achieved = gather_bytes / (gather_ms * 1e-3) / 1e9
        line = {
            "metric": "post-proc voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(shape, args.hops, args.mode), "tubes": n_tubes_for(shape, args.tubes),
                       "components": n_components, "labelled_voxels": labelled,
                       "l2": f"inputs larger than L2 ({ALGO_BYTES_PATH * V / world / 1e9:.1f} GB per GPU per step vs 126 MB)",
                       "sharding": "none" if world == 1 else (
                           f"Z-slabs x{world}; halo-run exchange + root all-gather "
                           + ("stored by the kernels into peer mailboxes over NVLink (release/acquire flags, no NCCL in a pass)"
                              if runner.transport == "peer" else "over NCCL send/recv + all-gather")),
                       "gather": ("split: stream phase on the main stream next to the labelling chain on a high-priority "
                                  "stream, then resolve") if (runner.split if world > 1 else split) else "fused, after the labelling",
                       "launch": "one CUDA graph per pass" if graphed else "eager launches"},
            "path_roofline": {"bytes_per_voxel": ALGO_BYTES_PATH, "achieved": ALGO_BYTES_PATH * V / world / (ms_per_step * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s",
                              "frac": ALGO_BYTES_PATH * V / world / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
            "roofline": {"kernel": f"{dominant}<half,int>", "bound": "hbm",
                         "timed": ("CUDA events around single launches on the same buffers right after the timed region"
                                   if (world > 1 or split) else "CUDA events around the kernel inside every timed step"), "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": measured_traffic(dominant, V / world),
                         "traffic_source": "profiles/r01_traffic.json (ncu --set full dram bytes per voxel x voxels per launch)",
                         "algorithmic_bytes": gather_bytes, "peak_source": peak_src,
                         "bytes_per_voxel": ALGO_BYTES_GATHER, "ms_per_launch": gather_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        if phases is not None:
            line["phases_ms_rank0"] = phases
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in an order that cannot wedge: drop the captured graph (it references the NCCL
        # communicator), drain the device, then destroy the group; a watchdog ends the process if the
        # NCCL teardown still blocks — the measurement has already been printed.
        import threading
        threading.Timer(30.0, lambda: os._exit(0)).start()
        runner.graph = None
        torch.cuda.synchronize(dev)
        dist.barrier()
        if hasattr(runner.comm, "close"):
            runner.comm.close()
        dist.destroy_process_group()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback; use --impl reference)")
        run_b200(args)


if __name__ == "__main__":
    main()
