"""CPU: host-side logic — the C-ABI library loads and exports every declared symbol, the cropper
mirror reproduces the reference's grid, patch_skoots rebinds the by-name imports, the product fails
loudly without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import skoots_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import skoots_b200._lib as L
    from skoots_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "skoots_b200.h")).read()
    declared = set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/skoots_b200.h but not exported"
    assert declared == set(L.SIGNATURES), "ctypes signature table out of sync with the header"
    assert L.load().skb_version() == 201


def test_argument_errors_do_not_need_a_gpu():
    import skoots_b200._lib as L
    lib = L.load()
    assert lib.skb_ccl_workspace_bytes(0, 4, 4, 16) == 0
    rc = lib.skb_ccl_label_sparse(None, 0, 4, 4, 4, 0, 2, 16, None, 0, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.skb_last_error()
    rc = lib.skb_assemble(None, 3, 1 << 20, 1 << 20, 4, L.f3((1, 1, 1)), 1, 1.0, L.i3((1, 1, 1)), L.i3((0, 0, 0)),
                          None, None, 0, None, 2, None)
    assert rc == -4  # SKB_E_RANGE: more than 2^31 voxels


def test_mailbox_geometry_is_host_arithmetic():
    """the peer transport's mailbox layout: every rank computes the same size from (world, capacities)."""
    import skoots_b200._lib as L
    lib = L.load()
    small = lib.skb_shard_mailbox_bytes(2, 1 << 10, 1 << 8, 1 << 7)
    big = lib.skb_shard_mailbox_bytes(8, 1 << 10, 1 << 8, 1 << 7)
    assert 0 < small < big
    # two copies of: two run buffers 3*(cap+1) ints, and world payloads of (2 + roots + 2*pairs) ints
    assert small >= 4 * (2 * 2 * 3 * ((1 << 10) + 1) + 2 * 2 * (2 + (1 << 8) + 2 * (1 << 7)))
    assert small % 256 == 0
    assert lib.skb_shard_mailbox_bytes(L.MAX_WORLD + 1, 16, 16, 16) == 0
    assert lib.skb_shard_mailbox_bytes(2, 0, 16, 16) == 0
    assert lib.skb_shard_begin(None, 2, 16, 16, 16, None) == -1 and b"NULL" in lib.skb_last_error()


def test_new_entry_points_validate_arguments_without_a_gpu():
    """argument errors of the round's new entry points come back as codes + text, nothing is enqueued."""
    import skoots_b200._lib as L
    lib = L.load()
    assert lib.skb_renumber_workspace_bytes(0, 8) == 0 and lib.skb_renumber_workspace_bytes(1 << 20, 1 << 10) > (1 << 20) // 8
    assert lib.skb_renumber(None, 2, 64, 8, None, 0, None, None, None, None) == -1 and b"NULL" in lib.skb_last_error()
    assert lib.skb_label_max(None, 5, 64, None, None) == -1
    assert lib.skb_unique_index_workspace_bytes(1) == 0 and lib.skb_unique_index_workspace_bytes(1000) >= 4000
    assert lib.skb_accuracies_from_iou(None, 0, 4, 0.5, None, None, None) == -1 and b"non-empty" in lib.skb_last_error()
    # slab geometry: Z, z_off, Zl multiples of 64; halo 1..64 planes
    assert lib.skb_shard_emit_runs(None, 8, 8, 100, 0, 64, 0, 12, None, 16, None, None) == -1
    assert b"multiples of 64" in lib.skb_last_error()
    assert lib.skb_assemble_stream(None, 3, 8, 8, 64, 0, 64, None, None, None, 2, 0, None) == -1
    assert lib.skb_assemble_stream(None, 3, 8, 8, 60, 0, 60, None, None, None, 2, 0, None) == -1
    assert lib.skb_peer_alloc(0, None) == -1 and lib.skb_peer_free(None) == 0 and lib.skb_peer_close(None) == 0


def test_no_cpu_fallback():
    import skoots_b200._lib as L
    from skoots_b200.lib.flood_fill import efficient_flood_fill
    from skoots_b200.lib.vector_to_embedding import vector_to_embedding
    with pytest.raises(L.SkootsB200Error):
        vector_to_embedding(torch.tensor((1, 1, 1)), torch.zeros((1, 3, 4, 4, 4)))
    with pytest.raises(L.SkootsB200Error):
        efficient_flood_fill(torch.zeros((4, 4, 4), dtype=torch.int16))


@pytest.mark.parametrize("dims,crop,ov", [((1, 70, 60, 24), [40, 40, 16], (5, 5, 2)), ((3, 2048, 300, 64), [500, 500, 50], (50, 50, 5)),
                                          ((1, 128, 128, 32), [500, 500, 50], (50, 50, 5)), ((1, 1100, 40, 24), [1000, 1000, 200], (0, 0, 0))])
def test_cropper_mirror_matches_oracle_grid(dims, crop, ov):
    from skoots_b200.lib.cropper import crops, get_total_num_crops
    img = torch.zeros(dims, dtype=torch.uint8)
    mine = list(crop)
    got = [tuple(o) for _, o in crops(img, mine, ov)]
    want = list(orc.crop_grid(dims[1:], crop, ov))
    assert got == want
    assert mine == orc.clamp_crop(dims[1:], crop)  # the caller's list is clamped in place, like the reference
    assert get_total_num_crops(img.shape, list(crop), ov) == len(want)
    first = next(iter(crops(img, list(crop), ov)))[0]
    assert first.shape == (1, dims[0]) + tuple(mine)


def test_cropper_guards_the_reference_hang():
    from skoots_b200.lib.cropper import crops
    with pytest.raises(ValueError):
        list(crops(torch.zeros((1, 64, 8, 64)), [500, 500, 50], (50, 50, 5)))


def test_patch_skoots_rebinds_reference_modules():
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    import skoots.lib.flood_fill
    import skoots.lib.vector_to_embedding
    import skoots_b200.lib.flood_fill as ff
    import skoots_b200.patch
    before = skoots.lib.flood_fill.efficient_flood_fill
    try:
        done = skoots_b200.patch.patch_skoots()
        assert ("skoots.lib.flood_fill", "efficient_flood_fill") in done
        assert skoots.lib.flood_fill.efficient_flood_fill is ff.efficient_flood_fill
        assert skoots.lib.vector_to_embedding.vector_to_embedding.__module__ == "skoots_b200.lib.vector_to_embedding"
        import skoots.validate.lib
        assert ("skoots.validate.lib", "mask_iou") in done
        assert skoots.validate.lib.mask_iou.__module__ == "skoots_b200.validate"
        # the crop grid: the defining module and the by-name copy inside the reference's flood fill (flood_fill.py:10)
        import skoots.lib.cropper
        assert skoots.lib.cropper.crops.__module__ == "skoots_b200.lib.cropper"
        assert skoots.lib.flood_fill.crops.__module__ == "skoots_b200.lib.cropper"
        assert ("skoots.lib.flood_fill", "crops") in done
    finally:
        skoots_b200.patch.unpatch_skoots()
    assert skoots.lib.flood_fill.efficient_flood_fill is before
    assert skoots.lib.flood_fill.crops.__module__ == "skoots.lib.cropper"


def test_patch_skoots_bug_compatible_binding():
    """patch_skoots(bug_compatible=True) binds the flood fill that reproduces the reference's multi-crop quirks."""
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    import skoots.lib.flood_fill
    import skoots_b200.patch
    try:
        skoots_b200.patch.patch_skoots(bug_compatible=True)
        fn = skoots.lib.flood_fill.efficient_flood_fill
        assert getattr(fn, "keywords", None) == {"reference_crops": True} and fn.__name__ == "efficient_flood_fill"
        import skoots.lib.skeleton
        bake = skoots.lib.skeleton.bake_skeleton  # the reference's own dispatch: Triton semantics for CUDA masks
        assert getattr(bake, "keywords", None) == {"triton_compat": "auto"} and bake.__name__ == "bake_skeleton"
    finally:
        skoots_b200.patch.unpatch_skoots()


def test_round2_entry_points_validate_arguments_without_a_gpu():
    """the entry points added in round 2 reject bad arguments with a code and a message; nothing is enqueued."""
    import skoots_b200._lib as L
    lib = L.load()
    f3, i3 = L.f3((60, 60, 12)), L.i3((500, 500, 50))
    # slab gather: NULL pointers, a slab that is not a multiple of 64 planes, a range that does not start on 256 voxels
    assert lib.skb_assemble_slab_ex(None, 3, 8, 8, 128, 0, 64, f3, 1, 1.0, None, None, None, None, 0, 0, 0, None, None, None, 12, None, 2, 0, 4096,
                                    None, None) == -1 and b"NULL" in lib.skb_last_error()
    assert lib.skb_assemble_slab_ex(1, 3, 8, 8, 128, 0, 60, f3, 1, 1.0, None, None, None, None, 0, 0, 0, 1, None, None, 12, 16, 2, 0, 3840,
                                    None, None) == -1 and b"multiples of 64" in lib.skb_last_error()
    assert lib.skb_assemble_slab_ex(1, 3, 8, 8, 128, 0, 64, f3, 1, 1.0, None, None, None, None, 0, 0, 0, 1, None, None, 12, 16, 2, 100, 256,
                                    None, None) == -1 and b"multiple of 256" in lib.skb_last_error()
    # N > 1 on an inner slab needs the neighbours' vector planes
    assert lib.skb_assemble_slab_ex(1, 3, 600, 600, 256, 64, 64, f3, 10, 1.0, i3, L.i3((50, 50, 5)), None, None, 0, 0, 0, 1, None, None, 64, 16, 1, 0,
                                    600 * 600 * 64, None, None) == -1 and b"vector halos" in lib.skb_last_error()
    assert lib.skb_assemble_planar(None, 3, 4, 64, 64, L.f3((60, 60)), None, None, 0, None, 2, None) == -1
    assert lib.skb_bake_skeletons(None, 2, 1, 8, 8, 8, None, None, None, 0, None, 0, L.f3((1, 1, 1)), 1, None, None, None, None, None) == -1
    assert lib.skb_elastic_resample(None, 2, 6, 6, L.f3((0.01, 0.05, 0.05)), None, None, 1, 8, 8, 8, None) == -1
    assert lib.skb_elastic_points(1, 2, 6, 6, L.f3((0.01, 0.05, 0.05)), None, 1, 5, 8, 8, 8, None, None) == -1
    assert lib.skb_shard_begin_pass(None, 2, 16, 16, 16, 64, None, None, None, None) == -1
    assert lib.skb_shard_push(None, 8, 8, 64, 16, None, None, None, 2, 0, 16, 16, 16, None, None) == -1


def test_compute_device_rules_without_a_gpu():
    """host tensors are staged through the GPU: without a CUDA device that is an error, never a CPU computation."""
    import skoots_b200._lib as L
    from skoots_b200.lib.morphology import binary_dilation
    from skoots_b200.validate import mask_iou
    if torch.cuda.is_available():
        pytest.skip("this test is about the box without a GPU")
    with pytest.raises(L.SkootsB200Error, match="no CPU fallback"):
        binary_dilation(torch.zeros((1, 1, 4, 4, 4)))
    with pytest.raises(L.SkootsB200Error, match="no CPU fallback"):
        mask_iou(torch.zeros((4, 4, 4), dtype=torch.int32), torch.zeros((4, 4, 4), dtype=torch.int32))
    with pytest.raises(L.SkootsB200Error):
        L.compute_device(torch.zeros(1), "not a tensor")


@pytest.mark.parametrize("module", ["sharded", "pipeline", "_lib", "validate", "patch"])
def test_every_private_method_a_class_calls_is_defined(module):
    """the sharded driver only runs on a GPU box: catch a method lost in an edit here, where there is none."""
    import ast
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "skoots_b200", module + ".py")
    src = open(path).read()
    for cls in [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef)]:
        defined = {f.name for f in ast.walk(cls) if isinstance(f, (ast.FunctionDef, ast.ClassDef))}
        assigned = set(re.findall(r"self\.(\w+)\s*=", ast.get_source_segment(src, cls)))
        called = set(re.findall(r"self\.(_\w+)\(", ast.get_source_segment(src, cls)))
        assert not (called - defined - assigned), (module, cls.name, sorted(called - defined - assigned))


def test_seam_test_of_the_reference_crops_path_equals_the_oracles():
    """row f3: the sum/product seam test (flood_fill.py:237-261) as presence-table lookups (device-agnostic torch) returns
    the oracle's pairs in the oracle's order, int16 wrap-around of large labels included."""
    import numpy as np
    from skoots_b200.lib.flood_fill import _adjacent_by_sum_product
    rng = np.random.default_rng(0)
    for trial in range(12):
        hi = (5, 200, 3000, 32767)[trial % 4]
        p0 = (rng.integers(0, hi + 1, size=(40, 50)) * (rng.random((40, 50)) < 0.3)).astype(np.int16)
        p1 = (rng.integers(0, hi + 1, size=(40, 50)) * (rng.random((40, 50)) < 0.3)).astype(np.int16)
        want = [(int(a), int(b)) for a, b in orc._adjacent_by_sum_product(p0, p1)]
        assert _adjacent_by_sum_product(torch.from_numpy(p0), torch.from_numpy(p1)) == want
    assert _adjacent_by_sum_product(torch.zeros((4, 4), dtype=torch.int16), torch.ones((4, 4), dtype=torch.int16)) == []


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """the boundary is a C ABI: include/skoots_b200.h compiles as C99 (-pedantic), and a plain C program links against
    the shared library and uses its host-side entry points (examples/c_abi_check.c; no GPU needed)."""
    import shutil
    import subprocess
    import skoots_b200._lib as L
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "skoots_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    L.load()
    libdir = os.path.dirname(L.LIB_PATH)
    exe = str(tmp_path / "c_abi_check")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "c_abi_check.c"), "-L", libdir, "-lskoots_b200", f"-Wl,-rpath,{libdir}", "-o", exe],
                   check=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "NULL argument -> -1" in run.stdout, run.stdout + run.stderr


def test_numa_local_is_best_effort_and_restores_the_affinity():
    """pinned buffers are allocated under `numa_local(device)`: without a GPU / NVML it must change nothing, say why,
    and leave the thread's CPU affinity as it found it."""
    import skoots_b200._lib as L
    before = os.sched_getaffinity(0)
    with L.numa_local("cuda:0") as info:
        inside = os.sched_getaffinity(0)
        assert info["bound"] in (False, True) and (info["bound"] or inside == before)
    assert os.sched_getaffinity(0) == before
    if not torch.cuda.is_available():
        assert info["bound"] is False and "why" in info
