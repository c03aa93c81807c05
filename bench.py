#!/usr/bin/env python
"""bench.py — post-processing throughput of the skeleton-embedding instance-assembly path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

One "step" = one pass of the hot path (connected-component labelling of the u8 skeleton mask +
fused vector->embedding->label gather) over the synthetic analytic-tube volume named in
`config.workload`.  `value` is voxels/s with inputs resident in HBM (CUDA events, max over
ranks); `e2e` is the same pass through the public host-buffer API with the host<->device copies
inside the timed region.  Every run also proves what it timed (`parity`):

  sharded_vs_unsharded   N > 1: every rank recomputes the whole volume unsharded on its own GPU and compares
                         its slab of the timed, sharded output bit for bit;
  sample_vs_oracle       the FULL-VOLUME output (not a re-run) restricted to a sample box against the CPU
                         reference run on that box (oracle/sample_check.py; canonical relabelling);
  e2e_vs_device          the host-buffer result equals the device-resident result.

See DESIGN.md §Measurement for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SCALE = (60, 60, 12)           # SKOOTS.VECTOR_SCALING default, skoots/config.py:144
ALGO_BYTES_PATH = 11.0         # u8 mask + 3 x fp16 vectors in, int32 label out
ALGO_BYTES_GATHER = 10.0       # dominant kernel (fused gather, or its stream phase): 6 B vectors in + 4 B labels out
EVAL_CROP, EVAL_OVERLAP = (500, 500, 50), (50, 50, 5)  # skoots/lib/eval.py:248-249
REFERENCE_ARM_BUDGET_S = 200.0  # whole `--impl reference --steps K --warmup W` run
CPU_BASELINE_BUDGET_S = 8.0     # one pass of the CPU path inside the b200 arm (the same box when K + W = 25)


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
    ap.add_argument("--e2e-out", default="int16", choices=["int16", "int32"],
                    help="dtype of the instance mask the host-buffer pass returns (the reference's is int16, eval.py:245)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU pass (also skips parity.sample_vs_oracle)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the unsharded recomputation at N > 1")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the eval()-mode N = 10 line and the density sweep")
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
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                rec = json.load(fh)
            if rec["kernel"].split("<")[0] != kernel:
                continue
            return rec["dram_bytes_per_voxel"] * voxels_per_launch, f"profiles/{name}"
        except Exception:
            continue
    return None, None


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
# what both arms agree on: the workload and the box of it the CPU path is timed on
# ----------------------------------------------------------------------------------------------
def mode_kwargs(mode):
    return dict(crop=EVAL_CROP, overlap=EVAL_OVERLAP) if mode == "eval" else dict(crop=None, overlap=(0, 0, 0))


def workload_name(shape, hops, mode="whole"):
    how = "whole-volume" if mode == "whole" else "eval() crop grid 500/500/50 ov 50/50/5,"
    return (f"synthetic analytic tubes {shape[0]}x{shape[1]}x{shape[2]} (BASELINE.json configs[2]), {how} "
            f"flood fill + vector_to_embedding(N={hops}) + index_skeleton_by_embed")


def cpu_sample_plan(shape, mode, hops, passes: int = 25):
    """the box of the workload the CPU path is timed on — the SAME box in the b200 arm's `cpu_baseline` and in
    every step of `--impl reference`: sized so that `passes` passes (the driver's 20 + 5) fit the reference arm's budget."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sample_check
    return sample_check.plan(shape, mode, hops, REFERENCE_ARM_BUDGET_S / max(1, passes))


def shared_config(shape, hops, mode, world):
    """identical in both arms (the driver compares them)."""
    plan = cpu_sample_plan(shape, mode, hops)
    V = shape[0] * shape[1] * shape[2]
    return {"workload": workload_name(shape, hops, mode), "tubes": n_tubes_for(shape, 0),
            "cpu_sample": plan["text"],
            "l2": f"inputs larger than L2 ({ALGO_BYTES_PATH * V / world / 1e9:.1f} GB per GPU per step vs 126 MB)"}, plan


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (oracle/_ref; the oracle port if the reference cannot be imported)
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sample_check
    from skoots_b200.synthetic import make_tube_volume
    torch.set_num_threads(os.cpu_count() or 1)
    full = tuple(int(v) for v in args.shape.split(","))
    config, plan = shared_config(full, args.hops, args.mode, max(1, args.gpus))
    if args.steps + args.warmup > 25:  # more passes than the box was sized for: shrink it (and say so)
        plan = sample_check.plan(full, args.mode, args.hops, REFERENCE_ARM_BUDGET_S / (args.steps + args.warmup))
        config["cpu_sample"] = plan["text"]
    R = plan["R"]
    tv = make_tube_volume(full, n_tubes_for(full, args.tubes), seed=0, z_range=(0, R[2]), xy_range=((0, R[0]), (0, R[1])),
                          want_mask=False, want_skeleton_dict=False)
    scale = torch.tensor(SCALE)
    kind = "port"
    for _ in range(args.warmup):
        kind = sample_check.run_cpu(tv.skeleton, tv.vectors, scale, args.hops, args.mode)[3]
    times = []
    for _ in range(args.steps):
        _, _, secs, kind = sample_check.run_cpu(tv.skeleton, tv.vectors, scale, args.hops, args.mode)
        times.append(secs)
    vox = R[0] * R[1] * R[2]
    value = vox * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "post-proc voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": plan["text"],
                         "what": ("the unmodified reference functions (skoots.lib.flood_fill.efficient_flood_fill, "
                                  "vector_to_embedding, index_skeleton_by_embed) via oracle/ref_runner.py" if kind == "reference"
                                  else "oracle/skoots_oracle.py (the reference tree is not importable here)")},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def sample_parity(out_full, mask_dev, vec_dev, shape, plan, hops, mode):
    """CPU path on the sample box R + comparison of the FULL-VOLUME device output on S.  Returns (detail, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sample_check
    R, S = plan["R"], plan["S"]
    torch.set_num_threads(os.cpu_count() or 1)
    smask = mask_dev[:R[0], :R[1], :R[2]].contiguous().cpu()
    svec = vec_dev[:, :R[0], :R[1], :R[2]].contiguous().cpu()
    scale = torch.tensor(SCALE)
    # warm numba's JIT and the thread pool (whole mode: a box this small has no room for eval()'s 50/50/5 margins)
    sample_check.run_cpu(smask[:96, :96, :32].contiguous(), svec[:, :96, :96, :32].contiguous(), scale, 1, "whole")
    want, labels, secs, kind = sample_check.run_cpu(smask, svec, scale, hops, mode)
    got = out_full[:S[0], :S[1], :S[2]].contiguous().cpu()
    detail = sample_check.compare(got, want, labels, S, shape)
    detail.update({"sample_box": list(R), "compared_box": list(S), "cpu_path": kind, "cpu_seconds": round(secs, 2)})
    return detail, secs, kind


def time_passes(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_extras(dev, mask, vec, shape, hbm_peak):
    """N = 1 only, after the headline measurement: (a) the configuration eval() actually runs — crop grid, N = 10, int16 —
    on the same full volume, its sample checked against the CPU reference; (b) a foreground-density sweep of the fused path."""
    from skoots_b200.pipeline import assemble_instances
    from skoots_b200.synthetic import make_tube_volume
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sample_check
    X, Y, Z = shape
    V = X * Y * Z
    scale = torch.tensor(SCALE)
    extras = {}
    # (a) eval() replay at full size
    out16 = torch.empty(shape, dtype=torch.int16, device=dev)
    ws = {"w": None}

    def eval_pass():
        assemble_instances(mask, vec, scale, N=10, crop=EVAL_CROP, overlap=EVAL_OVERLAP, out=out16, check=False, workspace=ws["w"])
    ms = time_passes(eval_pass, 3)
    plan = sample_check.plan(shape, "eval", 10, 8.0)
    detail, secs, kind = sample_parity(out16, mask, vec, shape, plan, 10, "eval")
    assert detail["ok"], f"eval()-mode N=10 output differs from the CPU reference on the sample: {detail}"
    Rv = plan["R"][0] * plan["R"][1] * plan["R"][2]
    extras["eval_N10"] = {"workload": workload_name(shape, 10, "eval"), "ms_per_step": ms, "voxels_per_s": V / (ms * 1e-3),
                          "path_roofline_frac": ALGO_BYTES_PATH * V / (ms * 1e-3) / 1e9 / hbm_peak,
                          "cpu_voxels_per_s": Rv / secs, "cpu_kind": kind, "cpu_sample": plan["text"],
                          "sample_vs_oracle": "bit-exact" if detail["ok"] else "MISMATCH", "parity_detail": detail}
    del out16
    # (b) density sweep on 1024x1024x256 (2.9 GB of inputs per pass: still far larger than L2)
    sshape = tuple(min(a, b) for a, b in zip((1024, 1024, 256), shape))
    sV = sshape[0] * sshape[1] * sshape[2]
    sweep = []
    for tubes, radius in ((1400, 4.0), (3200, 8.0), (8000, 12.0)):
        tubes = max(4, int(tubes * sV / (1024 * 1024 * 256)))
        tv = make_tube_volume(sshape, tubes, seed=1, device=dev, radius=radius, want_mask=False, want_skeleton_dict=False)
        out = torch.empty(sshape, dtype=torch.int32, device=dev)
        ms = time_passes(lambda: assemble_instances(tv.skeleton, tv.vectors, scale, N=1, out=out, check=False), 5)
        density = float((tv.vectors != 0).any(dim=0).float().mean().item())
        plan = sample_check.plan(sshape, "whole", 1, 1.0)
        detail, _, kind = sample_parity(out, tv.skeleton, tv.vectors, sshape, plan, 1, "whole")
        assert detail["ok"], f"density sweep ({density:.2f}): output differs from the CPU reference on the sample: {detail}"
        sweep.append({"nonzero_vector_fraction": round(density, 4), "skeleton_fraction": round(float(tv.skeleton.float().mean().item()), 4),
                      "tubes": tubes, "tube_radius": radius, "ms_per_step": ms, "voxels_per_s": sV / (ms * 1e-3),
                      "path_roofline_frac": ALGO_BYTES_PATH * sV / (ms * 1e-3) / 1e9 / hbm_peak,
                      "sample_vs_oracle": "bit-exact" if detail["ok"] else "MISMATCH"})
        del tv, out
    extras["density_sweep"] = {"shape": list(sshape), "rows": sweep}
    return extras


def run_b200(args):
    import torch.distributed as dist
    import skoots_b200._lib as L_
    from skoots_b200.lib.flood_fill import label_components
    from skoots_b200.pipeline import HostAssembler, assemble_instances, gather_instances
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
    config, plan = shared_config(shape, args.hops, args.mode, world)
    kw = mode_kwargs(args.mode)
    e2e_dtype = torch.int16 if (args.e2e_out == "int16" or args.mode == "eval") else torch.int32
    n_tubes = n_tubes_for(shape, args.tubes)

    if world > 1:
        from skoots_b200.sharded import PeerComm, ShardedAssembler, TorchDistComm
        transport = os.environ.get("SKB_TRANSPORT", "peer")
        make = lambda comm: ShardedAssembler(shape, world, rank, dev, scale=SCALE, hops=args.hops, comm=comm,
                                             split=bool(os.environ.get("SKB_SPLIT")),
                                             crop=kw["crop"], overlap=kw["overlap"],
                                             out_dtype=torch.int16 if args.mode == "eval" else torch.int32)
        try:
            runner = make(PeerComm() if transport == "peer" else TorchDistComm())
        except L_.SkootsB200Error as exc:  # collective failure (every rank raises): CUDA IPC is not usable on this box
            if transport != "peer":
                raise
            if rank == 0:
                print(f"bench.py: {exc}; falling back to the NCCL transport", file=sys.stderr, flush=True)
            runner = make(TorchDistComm())
        z0, z1 = runner.z_range
        tv = make_tube_volume(shape, n_tubes, seed=0, device=dev, z_range=(z0, z1), want_mask=False, want_skeleton_dict=False)
        runner.load(tv.skeleton, tv.vectors)
        del tv
        split = runner.split

        def step(timers=None):
            return runner.step(timers, check=False)
        launches_per_step = runner.launches_per_step
    else:
        tv = make_tube_volume(shape, n_tubes, seed=0, device=dev, want_mask=False, want_skeleton_dict=False)
        mask, vec = tv.skeleton, tv.vectors
        del tv
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
        # init, pack, tile, boundary, flatten, scan x2, rank, clear, publish, gather (split: stream + resolve) (+memsets);
        # large whole-volume N = 1 passes add the density probe and the second gather instantiation (one of the two returns at once)
        probed = (not split) and args.hops == 1 and args.mode == "whole" and V // 256 >= 65536 and Z % 64 == 0
        launches_per_step = 12 if split else (13 if probed else 11)

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
    else:
        runner.check_status()
        if os.environ.get("SKB_PHASE_TIMING"):
            phases = runner.profile_phases()

    sampler = ClockSampler(local) if rank == 0 else None
    timers = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graphed = False
    if world > 1 and not os.environ.get("SKB_NO_GRAPH") and runner.graphable:  # a pass with an NCCL exchange inside stays eager
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
    # the status word is sticky over the timed passes (graph replays included): a peer time-out, a list overflow or a gather
    # target beyond the halo in ANY timed step raises here, before a number is printed
    if world == 1:
        state["sparse"].check()
        n_components = state["sparse"].num_components
        labelled = int((out > 0).sum().item())
        timed_out = out
    else:
        n_components, labelled = runner.check()
        timed_out = runner.out.clone()  # this rank's slab of the timed, sharded result
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
    parity, parity_detail = {}, {}

    # ---- e2e: host buffers through the public API, result compared with the device-resident one ------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))
        if world == 1:
            with L_.numa_local(dev) as numa:  # pinned pages on the GPU's own NUMA node (matters when 8 GPUs copy at once)
                host_mask = torch.empty(shape, dtype=torch.uint8).pin_memory()
                host_vec = torch.empty((3,) + shape, dtype=torch.float16).pin_memory()
                host_out = torch.empty(shape, dtype=e2e_dtype).pin_memory()
            host_mask.copy_(mask)
            host_vec.copy_(vec)
            runner_h = HostAssembler(shape, dev, out_dtype=e2e_dtype)
            runner_h(host_mask, host_vec, scale, host_out, N=args.hops, **kw)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                runner_h(host_mask, host_vec, scale, host_out, N=args.hops, **kw)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / e2e_steps
            h2d, d2h = host_mask.numel() + host_vec.numel() * 2, host_out.numel() * host_out.element_size()
            e2e = {"value": V / dt, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": dt * 1e3, "steps": e2e_steps, "out_dtype": str(e2e_dtype).replace("torch.", ""),
                   "pcie_GBps_per_rank": [round((h2d + d2h) / dt / 1e9, 1)], "numa_rank0": numa,
                   "api": "skoots_b200.pipeline.HostAssembler"}
            same = True
            for x0 in range(0, X, max(1, X // 8)):  # compare on the device, piece by piece
                x1 = min(X, x0 + max(1, X // 8))
                same = same and bool(torch.equal(host_out[x0:x1].to(dev), timed_out[x0:x1].to(e2e_dtype)))
            del runner_h, host_mask, host_vec, host_out
        else:
            e2e = runner.e2e(e2e_steps, out_dtype=e2e_dtype)
            same = bool(torch.equal(runner.host_out.to(dev), timed_out.to(e2e_dtype)))
            ok = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            same = bool(ok.item())
        parity["e2e_vs_device"] = "bit-exact" if same else "MISMATCH"
        assert same, "the host-buffer (e2e) result differs from the device-resident result of the timed run"
        torch.cuda.empty_cache()

    # ---- N > 1: every rank recomputes the whole volume unsharded and compares its slab bit for bit --------------
    full_out = None
    if world > 1 and not args.no_parity:
        tv = make_tube_volume(shape, n_tubes, seed=0, device=dev, want_mask=False, want_skeleton_dict=False)
        mask, vec = tv.skeleton, tv.vectors
        del tv
        full_out = assemble_instances(mask, vec, scale, N=args.hops, out_dtype=timed_out.dtype, **kw)
        diff = int((full_out[:, :, z0:z1] != timed_out).sum().item())
        bad = torch.tensor([diff], device=dev, dtype=torch.int64)
        dist.all_reduce(bad)
        parity["sharded_vs_unsharded"] = "bit-exact" if int(bad.item()) == 0 else f"MISMATCH ({int(bad.item())} voxels)"
        parity_detail["sharded_vs_unsharded"] = (f"every rank recomputed the {X}x{Y}x{Z} volume unsharded on its own GPU and compared "
                                                 f"its slab of the timed {world}-rank result: {int(bad.item())} differing voxels over all ranks")
        assert int(bad.item()) == 0, f"rank {rank}: the sharded result differs from the unsharded one in {diff} voxels of its slab"
    elif world == 1:
        full_out = out
        parity["sharded_vs_unsharded"] = "n/a (1 GPU: the timed pass is the unsharded one)"

    # ---- rank 0: the FULL-VOLUME output on the sample box against the CPU path; the same pass is the CPU baseline ------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and full_out is not None:
        detail, secs, kind = sample_parity(full_out, mask, vec, shape, plan, args.hops, args.mode)
        parity["sample_vs_oracle"] = "bit-exact" if detail["ok"] else "MISMATCH"
        detail["note"] = ("the timed full-volume output restricted to compared_box vs the CPU path run on sample_box; equal up to the "
                          "canonical relabelling of component ids" + ("" if world == 1 else
                          "; at N > 1 the full-volume output is the unsharded recomputation that every rank's timed slab was just shown to equal"))
        parity_detail["sample_vs_oracle"] = detail
        assert detail["ok"], f"the full-volume GPU result differs from the CPU reference on the sample box: {detail}"
        if world == 1:
            R = plan["R"]
            cpu = {"value": R[0] * R[1] * R[2] / secs, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": plan["text"] + f" ({secs:.1f} s)",
                   "what": ("the unmodified reference functions via oracle/ref_runner.py (oracle/_ref)" if kind == "reference"
                            else "oracle/skoots_oracle.py (reference not importable here)")}

    extras = None
    if world == 1 and not args.no_extras and not args.no_cpu_baseline and args.mode == "whole" and args.hops == 1:
        del out, timed_out, full_out
        state["ws"] = state["sparse"] = None
        torch.cuda.empty_cache()
        extras = run_extras(dev, mask, vec, shape, hbm_peak)

    if rank == 0:
        gather_bytes = ALGO_BYTES_GATHER * V / world
        achieved = gather_bytes / (gather_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(dominant, V / world)
        # `config` stays what both arms agree on (workload, tube count, CPU sample box, input size); what only this arm
        # knows goes next to it
        details = {}
        details.update({"components": n_components, "labelled_voxels": labelled,
                       "sharding": "none" if world == 1 else (
                           f"Z-slabs x{world}; halo-run exchange + root all-gather "
                           + ("stored by the kernels into peer mailboxes over NVLink (release/acquire flags, no NCCL in a pass)"
                              if runner.transport == "peer" else "over NCCL send/recv + all-gather")),
                       "gather": ("split: stream phase on the main stream next to the labelling chain on a high-priority "
                                  "stream, then resolve") if split else "fused, after the labelling",
                       "launch": "one CUDA graph per pass" if graphed else "eager launches"})
        line = {
            "metric": "post-proc voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "details": details,
            "path_roofline": {"bytes_per_voxel": ALGO_BYTES_PATH, "achieved": ALGO_BYTES_PATH * V / world / (ms_per_step * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s",
                              "frac": ALGO_BYTES_PATH * V / world / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
            "roofline": {"kernel": f"{dominant}<half,int>", "bound": "hbm",
                         "timed": ("CUDA events around single launches on the same buffers right after the timed region"
                                   if (world > 1 or split) else "CUDA events around the kernel inside every timed step"), "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                         "traffic_source": f"{traffic_src} (ncu --set full dram bytes per voxel x voxels per launch)",
                         "algorithmic_bytes": gather_bytes, "peak_source": peak_src,
                         "bytes_per_voxel": ALGO_BYTES_GATHER, "ms_per_launch": gather_ms},
            "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "parity_detail": parity_detail,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        if extras is not None:
            line["extras"] = extras
        if phases is not None:
            line["phases_ms_rank0"] = phases
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in an order that cannot wedge: drop the captured graph (it references the NCCL communicator), drain the
        # device, close the peer mappings, destroy the group.  A watchdog turns a teardown that hangs into a NON-ZERO exit
        # (the line above has been printed, but the run must not look clean).
        import threading
        threading.Timer(60.0, lambda: os._exit(3)).start()
        runner.graph = None
        torch.cuda.synchronize(dev)
        dist.barrier()
        if hasattr(runner.comm, "close"):
            runner.comm.close()
        dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)  # past a completed teardown: skips interpreter-exit destructors of CUDA-IPC mappings that are already closed


def json_only_stdout():
    """The contract is ONE JSON line on stdout: anything a library printf()s to file descriptor 1 (NCCL's version banner
    under NCCL_DEBUG=VERSION/WARN) is sent to stderr, and Python's stdout keeps the original descriptor."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w", buffering=1)


def main():
    args = parse()
    json_only_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback; use --impl reference)")
        run_b200(args)


if __name__ == "__main__":
    main()
