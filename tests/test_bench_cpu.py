"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_json_line():
    out = _run(["--impl", "reference", "--width", "96", "--spp", "4", "--cpu-spp", "1", "--depth", "8", "--steps", "1",
                "--warmup", "0"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("Mrays/s") and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    out = _run(["--impl", "reference", "--width", "96", "--spp", "2", "--cpu-spp", "1", "--steps", "1", "--warmup", "0",
                "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible")
    out = _run(["--width", "96", "--spp", "1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline"])
    assert out.returncode != 0 and out.stdout.strip() == ""  # fails loudly: no JSON line, no CPU fallback


def test_reference_arm_reports_the_host_it_ran_on():
    out = _run(["--impl", "reference", "--width", "64", "--spp", "2", "--cpu-spp", "1", "--depth", "4", "--steps", "1",
                "--warmup", "0"])
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.strip()][0])
    host = d["cpu_baseline"]["host"]
    assert host["affinity"] == len(os.sched_getaffinity(0)) and "cgroup_cpus" in host and host["nproc"] >= 1
    assert d["cpu_baseline"]["cores"] <= max(host["affinity"], 1)
    assert "julia" in d and d["julia"]["found"] in (True, False)
    # the timing build: the same source rebuilt on this host; accepted only if it reproduces the portable build's bits
    assert "-O3 -march=native" in d["cpu_baseline"]["sample"] or "portable build" in d["cpu_baseline"]["sample"]
