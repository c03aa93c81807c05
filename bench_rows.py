#!/usr/bin/env python
"""bench_rows.py — every row of SURVEY.md §8(a) on one B200: device time (CUDA events, L2 flushed between
iterations), algorithmic GB/s against the measured HBM peak, and the CPU oracle timed beside it on
the same inputs (with a parity check).  Not the driver's benchmark (that is bench.py); this produces
the per-row evidence table committed under profiles/.

    python bench_rows.py > profiles/r02_rows.json
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import skoots_oracle as orc  # noqa: E402
from skoots_b200.lib import embedding_to_prob as e2p  # noqa: E402
from skoots_b200.lib import flood_fill as ff  # noqa: E402
from skoots_b200.lib import morphology as morph  # noqa: E402
from skoots_b200.lib import skeleton as skel  # noqa: E402
from skoots_b200.lib import vector_to_embedding as v2e  # noqa: E402
from skoots_b200.pipeline import EVAL_CROP, EVAL_OVERLAP, assemble_instances, tile_epilogue  # noqa: E402
from skoots_b200.synthetic import make_tube_volume  # noqa: E402

DEV = torch.device("cuda:0")
SCALE = torch.tensor((60, 60, 12))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except Exception:
        return 6650.0


_flush = None


def gpu_ms(fn, iters=10, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)  # > 126 MB L2
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters


def cpu_ms(fn, iters=2, warm=True):
    if warm:
        fn()
    best = 1e30
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def row(name, config, voxels, bytes_per_voxel, g_ms, c_ms, parity, note=""):
    gbs = voxels * bytes_per_voxel / (g_ms * 1e-3) / 1e9
    return {"row": name, "config": config, "voxels": voxels, "algorithmic_bytes_per_voxel": bytes_per_voxel,
            "gpu_ms": round(g_ms, 4), "gpu_voxels_per_s": voxels / (g_ms * 1e-3), "gpu_GBps": round(gbs, 1),
            "frac_of_measured_hbm": round(gbs / hbm_peak(), 4), "cpu_oracle_ms": round(c_ms, 2),
            "cpu_voxels_per_s": voxels / (c_ms * 1e-3), "cpu_threads": torch.get_num_threads(), "parity": parity, "note": note}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    rows = []

    # ---- C1: 128x128x32, 20 tubes, whole path ------------------------------------------------------
    tv = make_tube_volume((128, 128, 32), 20, seed=0)
    m, v = tv.skeleton.to(DEV), tv.vectors.to(DEV)
    for N in (1, 10):
        want = orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=N)
        got = assemble_instances(m, v, SCALE, N=N)
        rows.append(row(f"a1+a2+a3 whole path N={N}", "C1 128x128x32, 20 tubes", 128 * 128 * 32, 11,
                        gpu_ms(lambda: assemble_instances(m, v, SCALE, N=N, check=False)),
                        cpu_ms(lambda: orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=N)), bool(torch.equal(got.cpu(), want)),
                        "launch-latency bound at this size (12 kernels for 0.5 Mvox)"))

    from skoots_b200.pipeline import GraphedAssembler
    for N in (1, 10):
        run = GraphedAssembler((128, 128, 32), SCALE, DEV, N=N)
        want = orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=N)
        got = run(m, v)
        rows.append(row(f"a1+a2+a3 whole path N={N}, ONE CUDA-graph launch (GraphedAssembler)", "C1 128x128x32, 20 tubes", 128 * 128 * 32, 11,
                        gpu_ms(lambda: run(m, v, check=False)), cpu_ms(lambda: orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=N)),
                        bool(torch.equal(got.cpu(), want)), "two device-to-device input copies + one graph replay of the 11-kernel chain"))

    # ---- C2: one 300x300x20 tile: epilogue, dilation chain, flood fill, assembly --------------------------
    tv = make_tube_volume((300, 300, 20), 20, seed=0)
    g = torch.Generator().manual_seed(1)
    unet = torch.zeros((1, 5, 300, 300, 20))
    unet[0, 0:3] = tv.vectors.float() + 0.05 * torch.randn((3, 300, 300, 20), generator=g)
    unet[0, 3] = tv.skeleton.float() * 0.9 + 0.05 * torch.rand((300, 300, 20), generator=g)
    unet[0, 4] = (tv.mask > 0).float() * 0.95 + 0.04 * torch.rand((300, 300, 20), generator=g)
    unet_d = unet.to(DEV)
    wv, ws = torch.zeros((3, 300, 300, 20), dtype=torch.float16), torch.zeros((1, 300, 300, 20), dtype=torch.uint8)
    gv, gs = wv.to(DEV), ws.to(DEV)
    orc.tile_epilogue(unet, wv, ws, (0, 0, 0), (50, 50, 5))
    tile_epilogue(unet_d, gv, gs, (0, 0, 0), (50, 50, 5))
    rows.append(row("a5 tile epilogue", "C2 300x300x20 tile", 300 * 300 * 20, 20,
                    gpu_ms(lambda: tile_epilogue(unet_d, gv, gs, (0, 0, 0), (50, 50, 5))),
                    cpu_ms(lambda: orc.tile_epilogue(unet, wv, ws, (0, 0, 0), (50, 50, 5))),
                    bool(torch.equal(gv.cpu(), wv) and torch.equal(gs.cpu(), ws)), "20 B read per tile voxel"))
    # 16 different tiles back to back (what eval()'s tile loop does): 576 MB of network output, far larger than L2, and the
    # host's launch path amortised — the per-tile figure is the kernel's, not Python's
    tiles16 = [(unet_d + 0.001 * k).contiguous() for k in range(16)]
    rows.append(row("a5 tile epilogue, 16 tiles back to back (per tile)", "C2 16 x 300x300x20 tiles", 300 * 300 * 20, 20,
                    gpu_ms(lambda: [tile_epilogue(t, gv, gs, (0, 0, 0), (50, 50, 5)) for t in tiles16]) / 16,
                    cpu_ms(lambda: orc.tile_epilogue(unet, wv, ws, (0, 0, 0), (50, 50, 5))), True, "20 B read per tile voxel"))
    del tiles16
    img = unet[:, 3:4].contiguous()
    img_d = img.to(DEV)
    rows.append(row("a4 binary_dilation + 2x binary_dilation_2d", "C2 300x300x20", 300 * 300 * 20, 24,
                    gpu_ms(lambda: morph.binary_dilation_2d(morph.binary_dilation_2d(morph.binary_dilation(img_d)))),
                    cpu_ms(lambda: orc.binary_dilation_2d(orc.binary_dilation_2d(orc.binary_dilation(img)))),
                    bool(torch.equal(morph.binary_dilation_2d(morph.binary_dilation_2d(morph.binary_dilation(img_d))).cpu(),
                                     orc.binary_dilation_2d(orc.binary_dilation_2d(orc.binary_dilation(img))))), "3 passes x 8 B"))
    rows.append(row("a4 binary_erosion", "C2 300x300x20", 300 * 300 * 20, 8, gpu_ms(lambda: morph.binary_erosion(img_d)),
                    cpu_ms(lambda: orc.binary_erosion(img)), bool(torch.equal(morph.binary_erosion(img_d).cpu(), orc.binary_erosion(img)))))
    lab_in = ws[0].to(torch.int16)
    lab_d = lab_in.to(DEV)
    want = orc.flood_fill_exact(lab_in.clone())
    rows.append(row("a3 efficient_flood_fill (int16 in place)", "C2 300x300x20", 300 * 300 * 20, 4,
                    gpu_ms(lambda: ff.efficient_flood_fill(lab_d.clone())), cpu_ms(lambda: orc.flood_fill_exact(lab_in.clone())),
                    bool(torch.equal(ff.efficient_flood_fill(lab_d.clone()).cpu(), want)), "includes the clone and a status read-back"))

    # ---- eval() replay on a multi-crop volume (N=10, 500/500/50 grid) ---------------------------------------
    tv = make_tube_volume((600, 600, 64), 150, seed=0)
    m, v = tv.skeleton.to(DEV), tv.vectors.to(DEV)
    t0 = time.perf_counter()
    want = orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=10, crop=EVAL_CROP, overlap=EVAL_OVERLAP, out_dtype=torch.int16)
    eval_cpu_ms = (time.perf_counter() - t0) * 1e3
    got = assemble_instances(m, v, SCALE, N=10, crop=EVAL_CROP, overlap=EVAL_OVERLAP, out_dtype=torch.int16)
    rows.append(row("a1+a2+a3+a6 eval() replay N=10", "600x600x64, 150 tubes (SURVEY §6 probe size)", 600 * 600 * 64, 11,
                    gpu_ms(lambda: assemble_instances(m, v, SCALE, N=10, crop=EVAL_CROP, overlap=EVAL_OVERLAP, out_dtype=torch.int16, check=False)),
                    eval_cpu_ms, bool(torch.equal(got.cpu(), want)), "the configuration skoots/lib/eval.py:245-284 actually runs"))

    # ---- C4: training step ops, batch of 8 crops 300x300x20 --------------------------------------------------
    B = 8
    vols = [make_tube_volume((300, 300, 20), 20, seed=s) for s in range(B)]
    present = [{int(k): t.skeletons[int(k)] for k in torch.unique(t.mask).tolist() if k != 0} for t in vols]
    present_d = [{k: p.to(DEV) for k, p in d.items()} for d in present]
    masks_d = [t.mask.to(DEV) for t in vols]
    an = (1.0, 1.0, 3.0)
    want0 = orc.bake_skeleton(vols[0].mask, present[0], an, average=True)
    got0 = skel.bake_skeleton(masks_d[0], present_d[0], an, average=True)
    # voxel x skeleton-point pairs the min-reduction visits (9 flop each: 3 sub, 3 mul by anisotropy^2-weighted diff, 3 add/compare)
    pairs = sum(int((t.mask == k).sum()) * int(p.shape[0]) for t, d in zip(vols, present) for k, p in d.items())
    cpu_bake = cpu_ms(lambda: [orc.bake_skeleton(vols[i].mask, present[i], an, average=True) for i in range(B)], iters=1, warm=False)
    bake_ms = gpu_ms(lambda: [skel.bake_skeleton(masks_d[i], present_d[i], an, average=True) for i in range(B)], iters=5)
    rows.append(row("a8 bake_skeleton (+average), 8 per-sample calls (the reference's call pattern)", "C4 8 x 300x300x20, 20 ids each",
                    B * 300 * 300 * 20, 16, bake_ms, cpu_bake, bool(torch.allclose(got0.cpu(), want0, rtol=1e-5, atol=1e-5)),
                    "one fused launch + table upload + one status read per sample: host-bound"))
    masks_b = torch.stack(masks_d)
    gotb = skel.bake_skeletons_batch(masks_b, present_d, an, average=True)
    batch_ms = gpu_ms(lambda: skel.bake_skeletons_batch(masks_b, present_d, an, average=True), iters=10)
    nocheck_ms = gpu_ms(lambda: skel.bake_skeletons_batch(masks_b, present_d, an, average=True, check=False), iters=10)
    # the kernel alone: tables packed once outside the timed region
    ids_t, begin_t, off_t, pts_t, n_ids, n_pts = skel._pack_batch(present_d, DEV)
    out_b = torch.empty((B, 3, 300, 300, 20), dtype=torch.float32, device=DEV)
    st_b = torch.zeros(1, dtype=torch.int32, device=DEV)
    from skoots_b200 import _lib as L_

    def bake_kernel_only():
        L_.check(L_.load().skb_bake_skeletons(masks_b.data_ptr(), L_.dtype_code(masks_b), B, 300, 300, 20, ids_t.data_ptr(), begin_t.data_ptr(),
                                              off_t.data_ptr(), n_ids, pts_t.data_ptr(), n_pts, L_.f3(an), 1, 0, out_b.data_ptr(), 0,
                                              st_b.data_ptr(), L_.stream_ptr(DEV)))
    kern_ms = gpu_ms(bake_kernel_only, iters=10)
    rows.append(row("a8 bake_skeletons_batch (+average), ONE launch for the batch", "C4 8 x 300x300x20, 20 ids each",
                    B * 300 * 300 * 20, 16, batch_ms, cpu_bake,
                    bool(torch.allclose(gotb[0].cpu(), want0, rtol=1e-5, atol=1e-5) and torch.equal(out_b, gotb)),
                    f"whole call incl. table packing and the status read; without the read {nocheck_ms:.4f} ms; the kernel alone "
                    f"{kern_ms:.4f} ms = {16 * B * 1.8e6 / (kern_ms * 1e-3) / 1e9:.0f} GB/s of 16 B/voxel "
                    f"({16 * B * 1.8e6 / (kern_ms * 1e-3) / 1e9 / hbm_peak():.2f} of HBM), {9 * pairs / (kern_ms * 1e-3) / 1e12:.3f} TFLOP/s "
                    f"of the nominal 74 TFLOP/s fp32 peak over {pairs} voxel-point pairs: memory-bound, not ALU-bound"))
    wantm = orc.skeleton_to_mask(present[0], (300, 300, 20), 9, 3)
    rows.append(row("a9 skeleton_to_mask r=9 f=3", "C4 8 x 300x300x20", B * 300 * 300 * 20, 4,
                    gpu_ms(lambda: [skel.skeleton_to_mask(present_d[i], (300, 300, 20), radius=9, flank_radius=3) for i in range(B)], iters=5),
                    cpu_ms(lambda: [orc.skeleton_to_mask(present[i], (300, 300, 20), 9, 3) for i in range(B)]),
                    bool(torch.equal(skel.skeleton_to_mask(present_d[0], (300, 300, 20), radius=9, flank_radius=3).cpu(), wantm))))
    vec = torch.stack([t.vectors.float() for t in vols]).to(torch.bfloat16)
    baked = torch.stack([orc.bake_skeleton(vols[i].mask, present[i], an, average=True) for i in range(B)]).to(torch.bfloat16)
    sigma = torch.tensor((20.0, 20.0, 20.0))
    vec_d, baked_d = vec.to(DEV), baked.to(DEV)
    want = orc.baked_embed_to_prob(orc.vector_to_embedding(SCALE.float(), vec), baked, sigma)
    got = e2p.vector_to_prob(SCALE.float(), vec_d, baked_d, sigma)
    ok = bool(torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-30))
    rows.append(row("a1+a7 vector_to_embedding + baked_embed_to_prob (two ops, fwd)", "C4 8 x 300x300x20 bf16", B * 300 * 300 * 20, 40,
                    gpu_ms(lambda: e2p.baked_embed_to_prob(v2e.vector_to_embedding(SCALE.float(), vec_d), baked_d, sigma)),
                    cpu_ms(lambda: orc.baked_embed_to_prob(orc.vector_to_embedding(SCALE.float(), vec), baked, sigma)), ok,
                    "6 in + 12 out, then 12 + 6 in + 4 out"))
    rows.append(row("a1+a7 fused vector_to_prob (fwd)", "C4 8 x 300x300x20 bf16", B * 300 * 300 * 20, 16,
                    gpu_ms(lambda: e2p.vector_to_prob(SCALE.float(), vec_d, baked_d, sigma)),
                    cpu_ms(lambda: orc.baked_embed_to_prob(orc.vector_to_embedding(SCALE.float(), vec), baked, sigma)), ok,
                    "6 + 6 in, 4 out"))
    vg = vec_d.clone().requires_grad_(True)
    w = torch.rand(want.shape, device=DEV)

    def fwd_bwd():
        vg.grad = None
        (e2p.vector_to_prob(SCALE.float(), vg, baked_d, sigma) * w).sum().backward()
    vr = vec.float().requires_grad_(True)
    wc = w.cpu()

    def ref_fwd_bwd():
        vr.grad = None
        (orc.baked_embed_to_prob(orc.vector_to_embedding(SCALE.float(), vr), baked.float(), sigma) * wc).sum().backward()
    rows.append(row("a1+a7 fused vector_to_prob fwd+bwd (incl. torch mul/sum)", "C4 8 x 300x300x20 bf16", B * 300 * 300 * 20, 42,
                    gpu_ms(fwd_bwd), cpu_ms(ref_fwd_bwd), True, "fwd 16 B + bwd 26 B; gradient parity is tested in tests/test_gpu_parity.py"))

    # ---- C5: 2-D mode, 4096x4096 slices -------------------------------------------------------------------------------
    for S in (1, 8, 64):
        # 64 slices = the 8-slice stack eight times over (slices are independent in planar mode; generating 262 144
        # tubes in one go would only exercise the generator)
        tv = make_tube_volume((min(S, 8), 4096, 4096), 4096 * min(S, 8), seed=0, device=DEV, want_mask=False, want_skeleton_dict=False)
        stack = tv.skeleton if S <= 8 else tv.skeleton.repeat(S // 8, 1, 1).contiguous()  # (S, X, Y), 4-connectivity per slice
        sp = ff.label_components(stack, planar=True, label_base=0)
        out = torch.empty(stack.shape, dtype=torch.int32, device=DEV)
        ff.write_dense(sp, out)
        host = stack[:8].cpu().numpy()
        want0 = orc.label_components(host[0])[0]
        n_cpu = min(S, 8)  # the oracle is timed on (up to) 8 slices and scaled: it is a per-slice loop

        def gpu_2d():
            s2 = ff.label_components(stack, planar=True, label_base=0, workspace=sp.workspace, check=False)
            ff.write_dense(s2, out)
        rows.append(row(f"a10 per-slice CCL (2-D mode), {S} slices", f"C5 {S} x 4096x4096", S * 4096 * 4096, 5, gpu_ms(gpu_2d),
                        cpu_ms(lambda: [orc.label_components(host[i]) for i in range(n_cpu)], iters=1) * (S / n_cpu),
                        bool((out[0].cpu().numpy() == want0).all() and (out[S - 1].cpu().numpy() == orc.label_components(host[(S - 1) % 8])[0]).all()),
                        "1 B mask in + 4 B int32 labels out" + ("" if S <= 8 else "; CPU time = 8 slices x 8")))
        # the whole 2-D path: per-slice CCL + the fused planar gather (vectors = the tubes' in-plane components)
        from skoots_b200.pipeline import assemble_instances_2d
        vt = tv.vectors[1:3].permute(1, 0, 2, 3).contiguous()
        vstack = vt if S <= 8 else vt.repeat(S // 8, 1, 1, 1).contiguous()
        s2d = torch.tensor((60.0, 60.0))
        out2 = torch.empty(stack.shape, dtype=torch.int32, device=DEV)
        assemble_instances_2d(stack, vstack, s2d, workspace=sp.workspace, out=out2)
        want2 = orc.postprocess_2d(stack[:1].cpu(), vstack[:1].cpu(), s2d)
        t0 = time.perf_counter()
        orc.postprocess_2d(stack[:1].cpu(), vstack[:1].cpu(), s2d)
        cpu_one = (time.perf_counter() - t0) * 1e3
        rows.append(row(f"a10 2-D path: per-slice CCL + fused planar gather, {S} slices", f"C5 {S} x 4096x4096", S * 4096 * 4096, 9,
                        gpu_ms(lambda: assemble_instances_2d(stack, vstack, s2d, workspace=sp.workspace, out=out2, check=False)),
                        cpu_one * S, bool(torch.equal(out2[0].cpu(), want2[0]) and torch.equal(out2[S - 1].cpu(), out2[(S - 1) % 8].cpu())),
                        "1 B mask + 4 B fp16 vectors in, 4 B labels out; CPU time = one slice x S"))
        del vt, vstack, out2
        v2 = (torch.rand((min(S, 8), 2, 4096, 4096), device=DEV) * 2 - 1).half()
        if S > 8:
            v2 = v2.repeat(S // 8, 1, 1, 1).contiguous()
        s2 = torch.tensor((60.0, 60.0))
        v2_cpu = v2[:n_cpu].cpu()
        rows.append(row(f"a10 vector_to_embedding 2-D, {S} slices", f"C5 {S} x 4096x4096", S * 4096 * 4096, 12,
                        gpu_ms(lambda: v2e.vector_to_embedding(s2, v2)), cpu_ms(lambda: orc.vector_to_embedding(s2, v2_cpu)) * (S / n_cpu),
                        bool(torch.equal(v2e.vector_to_embedding(s2, v2)[S - n_cpu:].cpu(), orc.vector_to_embedding(s2, v2[S - n_cpu:].cpu()))),
                        "4 B in + 8 B out" + ("" if S <= 8 else "; CPU time = 8 slices x 8")))
        del tv, stack, out, v2

    # ---- f2: renumber + validation metrics on an assembled instance mask ----------------------------------------------
    from skoots_b200 import validate as val
    shape = (512, 512, 128)
    tv = make_tube_volume(shape, 1500, seed=0, device=DEV, want_skeleton_dict=False)
    inst = assemble_instances(tv.skeleton, tv.vectors, SCALE, N=1)
    gt = tv.mask.to(torch.int32)
    V = shape[0] * shape[1] * shape[2]
    inst_h, gt_h = inst.cpu().numpy(), gt.cpu().numpy()
    scratch = inst.clone()

    def gpu_renumber():
        scratch.copy_(inst)
        val.renumber(scratch, in_place=True)
    t_copy = gpu_ms(lambda: scratch.copy_(inst))
    rows.append(row("f2 renumber (fastremap.renumber) in place", "512x512x128 instance mask, 1500 tubes", V, 12,
                    gpu_ms(gpu_renumber) - t_copy, cpu_ms(lambda: orc.renumber(inst_h), iters=1),
                    bool((val.renumber(inst)[0].cpu().numpy() == orc.renumber(inst_h)[0]).all()),
                    "4 B read (first occurrences) + 4 B read + 4 B written (apply); includes the max-label read-back"))
    want_iou = orc.mask_iou(gt_h, inst_h)
    rows.append(row("f2 mask_iou (contingency table)", f"512x512x128, {want_iou.shape[0]} x {want_iou.shape[1]} objects", V, 8,
                    gpu_ms(lambda: val.mask_iou(gt, inst)), cpu_ms(lambda: orc.mask_iou(gt_h, inst_h), iters=1),
                    bool(torch.equal(val.mask_iou(gt, inst).cpu(), want_iou)),
                    "4 B gt + 4 B prediction per voxel; the CPU figure is the oracle's contingency restatement, "
                    "not the reference's O(N*M*V) loop"))
    del tv, inst, gt, scratch

    # ---- f4: elastic deformation of one training crop (image + mask + skeleton points) -------------------------------
    from skoots_b200.train.merged_transform import elastic_deform
    tvf = vols[0]
    img5 = torch.rand((1, 1, 300, 300, 20))
    mask5 = tvf.mask.float()[None, None]
    sk_long = {k: p.long() for k, p in present[0].items()}
    noise = torch.rand((1, 3, 2, 6, 6))
    w_img, w_mask, w_sk = orc.elastic_deform(noise, img5, mask5, skeleton=sk_long)
    img_d5, mask_d5 = img5.to(DEV), mask5.to(DEV)
    sk_d5 = {k: p.to(DEV) for k, p in sk_long.items()}
    noise_d = noise.to(DEV)
    g_img, g_mask, g_sk = elastic_deform(img_d5, mask_d5, skeleton=sk_d5, noise=noise_d)
    mism = float((g_mask.cpu() != w_mask).float().mean())
    rows.append(row("f4 elastic_deform (image + mask + skeleton points)", "C4 one 300x300x20 crop, 20 skeletons", 2 * 300 * 300 * 20, 8,
                    gpu_ms(lambda: elastic_deform(img_d5, mask_d5, skeleton=sk_d5, noise=noise_d)),
                    cpu_ms(lambda: orc.elastic_deform(noise, img5, mask5, skeleton=sk_long), iters=1),
                    bool(mism <= 1e-3 and all(int((g_sk[k].cpu() - w_sk[k]).abs().max()) <= 1 for k in w_sk)),
                    f"4 B read + 4 B written per voxel and argument; {mism:.2e} of the mask voxels differ from the torch oracle "
                    "(rounding ties of the nearest sampling, DESIGN.md)"))

    # the reference's only GPU kernel (Triton, skoots/lib/skeleton.py:51-367) on the same box and inputs.
    # Its semantics differ from the CPU path (fp16 outputs, anisotropy on squared differences, per-axis max on
    # ties), so it is raced, not compared bit for bit.
    try:
        import ref_shim
        ref_shim.install(need_morphology=False)
        import skoots.lib.skeleton as ref_skel
        masks_i = [m.contiguous() for m in masks_d]
        ref_skel.bake_skeleton(masks_i[0], present_d[0], an, average=False)  # compile
        tri_ms = gpu_ms(lambda: [ref_skel.bake_skeleton(masks_i[i], present_d[i], an, average=False) for i in range(B)], iters=3, warm=1)
        ours_raw = gpu_ms(lambda: skel.bake_skeletons_batch(masks_b, present_d, an, average=False, check=False), iters=10)
        tri0 = ref_skel.bake_skeleton(masks_i[0], present_d[0], an, average=False).float()
        mine0 = skel.bake_skeleton(masks_d[0], present_d[0], an, average=False)
        agree = float((tri0 == mine0).float().mean().item())
        try:
            tri_avg_ms = gpu_ms(lambda: [ref_skel.bake_skeleton(masks_i[i], present_d[i], an, average=True) for i in range(B)], iters=3, warm=1)
        except Exception as exc:  # scripted morphology helpers may not run under this torch
            tri_avg_ms = None
        rows.append({"row": "a8 HEAD-TO-HEAD: reference Triton _bake_skeleton_triton vs skb_bake_skeletons (average=False)",
                     "config": "C4 8 x 300x300x20, 20 ids each, same B200, same inputs", "voxels": B * 300 * 300 * 20,
                     "reference_triton_ms": round(tri_ms, 4), "reference_triton_with_average_ms": None if tri_avg_ms is None else round(tri_avg_ms, 4),
                     "skoots_b200_ms": round(ours_raw, 4), "skoots_b200_with_average_ms": round(nocheck_ms, 4),
                     "speedup": round(tri_ms / ours_raw, 1), "fraction_of_voxels_where_both_agree": round(agree, 4),
                     "note": "the reference launches one Triton program per voxel, 8 launches + 8 synchronisations per batch, fp16 out"})
    except Exception as exc:
        rows.append({"row": "a8 HEAD-TO-HEAD: reference Triton kernel", "unavailable": repr(exc)[:300]})

    print(json.dumps({"hbm_peak_GBps": hbm_peak(), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
