"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (the CPU port of the reference's
kernels on all host cores) prints ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints; the GPU arm
fails loudly when there is no device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--qubits", "16", "--steps", "2", "--warmup", "1", "--ref-gates-per-step", "3"],
             env={"OMP_NUM_THREADS": "1"})          # torchrun sets this; the arm must use all cores anyway
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "gates/s" and d["unit"] == "gates/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_only_rank_zero_works():
    p = _run(["--impl", "reference", "--gpus", "2", "--qubits", "14", "--steps", "1", "--warmup", "1"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run(["--qubits", "12", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], env={"CUDA_VISIBLE_DEVICES": ""})
    assert p.returncode != 0
    assert p.stdout.strip() == ""            # no line, no silent CPU result
