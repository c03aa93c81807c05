"""GPU: bench.py keeps the driver's JSON contract (keys, units, the roofline / cpu_baseline / e2e objects) — run on a
small volume so that it takes seconds."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--shape", "256,256,128", "--steps", "3", "--warmup", "3", *extra],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    return json.loads(lines[0])


def test_bench_line_has_the_contract_keys():
    d = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "voxels/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] == pytest.approx(256 * 256 * 128 / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert "workload" in d["config"] and "model" not in d["config"] and "components" in d["details"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0
    assert d["config"]["cpu_sample"] in c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 256 * 256 * 128 * 7 and e["d2h_bytes_per_step"] == 256 * 256 * 128 * 2
    assert d["gpu_launches"] > 0 and e["value"] < d["value"]
    p = d["parity"]
    assert p["sample_vs_oracle"] == "bit-exact" and p["e2e_vs_device"] == "bit-exact"
    assert d["parity_detail"]["sample_vs_oracle"]["labelled_compared"] > 0
    x = d["extras"]
    assert x["eval_N10"]["sample_vs_oracle"] == "bit-exact" and x["eval_N10"]["voxels_per_s"] > 0
    assert len(x["density_sweep"]["rows"]) == 3 and all(r["sample_vs_oracle"] == "bit-exact" for r in x["density_sweep"]["rows"])
    fr = [r["nonzero_vector_fraction"] for r in x["density_sweep"]["rows"]]
    assert fr[0] < fr[1] < fr[2]


def test_bench_eval_mode_line():
    d = _run("--mode", "eval", "--hops", "10", "--no-extras")
    assert d["parity"]["sample_vs_oracle"] == "bit-exact" and d["parity"]["e2e_vs_device"] == "bit-exact"
    assert "eval()" in d["config"]["workload"] and d["e2e"]["out_dtype"] == "int16"
