"""CPU: the parts of bench.py that need no GPU -- the reference arm (`--impl reference`, the reference's own CPU code out of
oracle/_ref, or the oracle port when that is absent) prints the contract's JSON line; under torchrun only rank 0 works; the
GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.pop("WORLD_SIZE", None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, cwd=str(ROOT), env=e, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-crop", "256", "--planes", "3", "--size", "1024"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "crop" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # value is the sample's pixels over its seconds
    assert abs(d["value"] - 3 * 256 * 256 / 1e6 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_do_no_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-crop", "256", "--planes", "1"],
             env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
