"""bench.py's contract where it can be checked without a GPU: the CPU arm (`--impl reference` = the oracle port of the
reference path on the host cores) prints ONE JSON line with the keys the driver reads, its `config` is the product arm's
`config` for the same workload, ranks other than 0 stay silent, and the product arm fails loudly (no CPU fallback) when
there is no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_bench(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.fixture(scope="module")
def reference_line():
    r = run_bench("--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys(reference_line):
    d = reference_line
    assert d["impl"] == "reference"
    assert d["metric"] == "2nn_dist_evals_per_s" and d["unit"] == "dist-evals/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "queries" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_describe_the_workload_identically(reference_line):
    import bench
    for name, cfg in bench.WORKLOADS.items():
        c = bench.workload_config(cfg)
        assert c["workload"] == cfg["desc"] and "engine" not in c and "sample" not in c, name     # how an arm ran goes under `run`
    assert reference_line["config"] == bench.workload_config(bench.WORKLOADS["cfg2"])
    # the CPU arm's sample is at least 1/16 of the workload in both dimensions
    assert "20000 of 20000 queries" in reference_line["cpu_baseline"]["sample"]


def test_reference_arm_runs_on_rank_zero_only():
    r = run_bench("--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "0", "--gpus", "2",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the product arm would run")
    r = run_bench("--workload", "cfg2", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
