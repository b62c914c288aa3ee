"""bench.py contract checks that need no GPU: the reference arm prints the contract's JSON line (bounded CPU sample of the
same workload, oracle port), and the product arm fails loudly without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--syn", "200000", "--hidden", "2000", "--events", "50000"]


def _run(args, **kw):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True, timeout=300, **kw)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"] + SMALL)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "synaptic events/sec" and d["unit"] == "events/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # one step = events_per_step events at the printed rate
    assert abs(d["config"]["events_per_step"] / (d["ms_per_step"] * 1e-3) - d["value"]) <= 0.05 * d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"] + SMALL, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0", "--skip-cpu"] + SMALL)
    assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
