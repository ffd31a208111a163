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
           "--warmup", "1", "--workload", "poisson27:64", "--cpu-planes", "8", "--cpu-its", "2"] + list(extra)
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
    assert d["config"]["workload"] == "poisson27:64" and d["data"] == "synthetic"
    # the same config dict as the b200 arm prints (bench.config_of): what `same_config` compares
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_of("poisson27:64", d["config"]["iterations"])
    assert d["config"]["n"] == 64 ** 3 and d["config"]["nnz"] == (3 * 64 - 2) ** 3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "s"
    assert "8 z-planes of the real poisson27 64^3 operator (32768 rows" in cb["sample"]
    assert cb["value"] == d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_every_core_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must not run single-threaded
    (round 1: the N >= 2 reference legs timed out)"""
    r = run(env=dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == os.cpu_count()


def test_reference_arm_other_ranks_stay_silent():
    r = run(env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
