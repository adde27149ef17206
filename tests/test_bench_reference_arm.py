"""The reference arm of bench.py (`--impl reference`: the oracle port on the host cores, the tier's CPU baseline) runs without a GPU, so its
contract is checked here: one JSON line with the base contract's keys, all host cores in use even when the launcher exports
OMP_NUM_THREADS=1 (torchrun does; round 1 measured the arm on one thread at N >= 2), and under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "0", "--cpu-sample", "5e4", "--cpu-adj-sample", "1e4"]


def json_lines(text):
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            out.append(json.loads(line))
    return out


def check_line(d, n_gpus):
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"] == "loglik+gibbs_sweep_events_per_s" and d["unit"] == "events/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_prints_one_contract_line_and_uses_every_core():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + SMALL, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = json_lines(r.stdout)
    assert len(lines) == 1
    check_line(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"] + SMALL
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = json_lines(r.stdout)
    assert len(lines) == 1
    check_line(lines[0], 2)
