"""bench.py's reference arm runs on the CPU: check its JSON line against the
contract (one line on stdout; metric / unit / config equal to the b200 arm's;
cpu_baseline and e2e objects), on a tiny sample so it takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
           "--warmup", "1", "--cpu-sample-n", "40", "--cpu-sample-its", "20"] + list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_prints_one_contract_line():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pcg_time_to_1e-10" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["scaling"] == "strong" and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "poisson27:512" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "s" and "poisson27 40^3" in cb["sample"]
    assert cb["value"] == d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = run(env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
