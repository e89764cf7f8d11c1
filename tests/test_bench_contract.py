"""bench.py contract of the reference (CPU) arm: one JSON line with the agreed keys; under torchrun only
rank 0 works and prints.  Runs the small README workload so that it takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                           "--steps", "1", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = run({"RANK": "0", "WORLD_SIZE": "1"})
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "prepare_state_throughput" and d["unit"] == "states/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 2
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 1.0) < 1e-9
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    res = run({"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0, res.stderr
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_batch_workload_at_n_gt_1():
    """--gpus N > 1 selects the sharded batch workload (config 5); its CPU arm runs one oracle process per
    host core on rank 0."""
    env = dict(os.environ)
    env.update({"RANK": "0", "WORLD_SIZE": "2"})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["scaling"] == "strong" and d["config"]["batch"] == 4096
    assert d["config"]["n_qubits"] == 12 and d["config"]["layers"] == 10 and d["config"]["sweeps"] == 20
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
