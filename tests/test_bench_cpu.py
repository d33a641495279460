"""bench.py contract on a machine without a GPU: the reference arm (the reference's CPU algorithm -- sparse direct LU -- through the
oracle, the one place besides tests/ and smoke() that may execute oracle/) prints one JSON line with the keys the driver reads, and
the product arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-grid", "128", "--ref-fit", "64,128,192")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    z = json.loads(lines[0])
    assert z["impl"] == "reference" and z["metric"] == "solves_per_sec_4096x4096_TM_to_1e-10" and z["unit"] == "solves/s"
    assert z["n_gpus"] == 1 and z["steps"] == 1 and z["warmup"] == 0 and z["higher_is_better"] is True
    assert z["value"] > 0 and z["ms_per_step"] > 0 and z["vs_baseline"] is None and z["dtype"] == "c128" and z["data"] == "synthetic"
    assert "workload" in z["config"] and "model" not in z["config"]
    # the arm says which grid it really solved and how it extrapolated (round-1 finding: a 512^2 run labelled 4096^2)
    assert z["config"]["grid_solved_by_this_arm"] == [128, 128] and z["config"]["extrapolated_to"] == [4096, 4096]
    cb = z["cpu_baseline"]
    assert len(cb["fit_samples"]) == 3 and 0.8 < cb["fit_exponent"] < 2.5
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == z["value"] and cb["sample"]
    e = z["e2e"]
    assert e["value"] == z["value"] and e["unit"] == z["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = _run("--steps", "1", "--warmup", "0", "--grid", "256", "--no-cpu-baseline", timeout=300)
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{") and '"value"' in ln]
