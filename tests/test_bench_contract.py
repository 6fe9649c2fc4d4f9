"""bench.py's reference arm (the only arm that runs without a GPU) prints exactly one JSON line with the contract's keys;
under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(args, env=None):
    r = subprocess.run([sys.executable, *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run(["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1", "--frames", "4"])
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "1024x1024" in d["config"]["workload"]


def test_reference_arm_under_torchrun_only_rank0_prints():
    lines = run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29588",
                 "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--frames", "2"])
    js = [l for l in lines if l.startswith("{")]
    assert len(js) == 1, lines
    d = json.loads(js[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and "tiled" in d["config"]["workload"]
