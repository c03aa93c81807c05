"""CPU: `bench.py --impl reference` times the reference's own CPU path — the unmodified `skoots.lib` functions through
oracle/ref_runner.py (from /root/reference here, from the oracle/_ref copy on the GPU box) — and prints the driver's JSON
line.  A small --shape keeps this to seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    env = dict(os.environ)
    env.pop("PYTORCH_JIT", None)  # the timed path is the reference's stock one, torch.jit.script decorators active
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--shape", "256,192,64", *extra], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])


def test_reference_arm_line():
    import ref_shim
    d = _run()
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert "workload" in d["config"] and d["config"]["cpu_sample"] == d["cpu_baseline"]["sample"]
    assert d["value"] == 256 * 192 * 64 / (d["ms_per_step"] * 1e-3) or abs(d["value"] * d["ms_per_step"] * 1e-3 / (256 * 192 * 64) - 1) < 1e-9


def test_reference_arm_eval_mode():
    d = _run("--mode", "eval", "--hops", "3")
    assert d["impl"] == "reference" and d["value"] > 0 and "eval()" in d["config"]["workload"]


def test_both_arms_declare_the_same_workload_and_sample():
    """the driver compares the arms' `config`: workload, tube count, the CPU sample box and the input-size note are the
    same function of (shape, mode, hops, N) in both arms."""
    sys.path.insert(0, ROOT)
    import bench
    a, plan_a = bench.shared_config((2048, 2048, 512), 1, "whole", 1)
    b, plan_b = bench.shared_config((2048, 2048, 512), 1, "whole", 1)
    assert a == b and plan_a == plan_b
    assert plan_a["R"] == (960, 960, 160) and plan_a["S"] == (832, 832, 128)   # fits one of the reference's 1000x1000x200 flood-fill crops
    e, plan_e = bench.shared_config((2048, 2048, 512), 10, "eval", 1)
    assert plan_e["R"][2] == 90 and plan_e["R"][0] in (500, 900)               # a box whose crop grid is the full volume's grid
