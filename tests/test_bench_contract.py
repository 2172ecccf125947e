"""CPU tier: the parts of bench.py's contract that need no GPU -- the reference arm prints one well-formed JSON
line (and only rank 0 prints under a multi-rank launch), and the `ours` arm refuses to run without a CUDA device
instead of falling back to the host."""
import json
import os
import subprocess
import sys

import pytest

from .conftest import ROOT


def run_bench(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--samples", str(1 << 22)])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "rotate_cfg1"


def test_reference_arm_only_rank0_prints():
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ours_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench(["--steps", "1", "--warmup", "3", "--samples", "4096"])
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)


def test_config_passes_and_arms_agree():
    """Every extra configuration the `ours` arm measures names a known workload, and both arms build the `config` object
    from the same function, so that the driver's `same_config` check compares like with like."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("zc_bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    for wl, pm in b.CONFIG_PASSES:
        assert wl in b.WORKLOADS and pm in ("sweep", "random", "nco"), (wl, pm)
        kind, nper, bytes_per, opts = b.WORKLOADS[wl]
        assert nper >= 1 << 28 and bytes_per in (4, 6, 8, 12, 16, 20)
        assert b.core_name(kind, opts)
    kind, nper, bytes_per, opts = b.WORKLOADS["rotate_cfg1"]
    cfg = b.config_for("rotate_cfg1", kind, opts, nper, "sweep", bytes_per)
    assert cfg["samples_per_gpu_per_step"] == 1 << 30 and cfg["workload"] == "rotate_cfg1"
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--samples", str(1 << 22)])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert set(d["config"]) == set(cfg)                       # same keys; the values differ only through --samples here
