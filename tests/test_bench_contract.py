"""CPU test of bench.py's reference arm (the only arm that runs without a GPU): one JSON line with the contract's
keys, rank 0 alone prints under a multi-rank launch, and the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    out = run(["--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "sclens_cells_per_s" and j["unit"] == "cells/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1 and j["warmup"] == 0
    assert j["steps_requested"] == 1 and j["warmup_requested"] == 0
    assert j["config"]["workload"].startswith("small") and j["data"] == "synthetic" and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "measured once at FULL size" in cb["sample"]
    # the stages are measured, not modelled: seconds per stage at the full shape of the workload, measured once
    assert cb["frac"] == 1 and cb["measured_at"] == [2000, 3000] and cb["eig_n"] == 2000
    assert set(cb["stage_seconds"]) == {"normalise", "gram_f64", "eigen_vectors", "corr"} and min(cb["stage_seconds"].values()) > 0
    assert j["e2e"] == {"value": j["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = run(["--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0", "--gpus", "2"],
              env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = run(["--workload", "small", "--steps", "1", "--warmup", "1"])
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
