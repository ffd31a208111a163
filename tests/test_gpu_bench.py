"""bench.py on a B200, small workload: exactly one JSON line with every key of
the measurement contract (roofline with the measured peak, e2e with the host
copies, cpu_baseline, gpu_launches, clocks) and a solve that really converged."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "poisson27:96",
           "--steps", "2", "--warmup", "3", "--cpu-sample-n", "48", "--cpu-sample-its", "30"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["metric"] == "pcg_time_to_1e-10" and d["unit"] == "s" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["higher_is_better"] is False and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "poisson27:96" and d["config"]["n"] == 96 ** 3
    assert abs(d["ms_per_step"] - 1e3 * d["value"]) <= 1e-9 * d["ms_per_step"]
    it = d["config"]["iterations"]
    assert 150 < it < 300 and d["pcg"]["true_relres"] <= 1e-10      # ~2.27 N
    assert d["gpu_launches"] >= 3 * 2 * it                              # 3 kernels per iteration, 2 steps
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000
    assert rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["algorithmic_bytes_per_launch"] == 12 * d["config"]["nnz_local"] + 4 * (96 ** 3 + 1) + 16 * 96 ** 3
    assert rf["stored_bytes_per_launch"] < rf["algorithmic_bytes_per_launch"]   # index compression is on
    e = d["e2e"]
    assert e["unit"] == "s" and e["value"] >= d["value"] * 0.9
    assert e["h2d_bytes_per_step"] == 2 * 8 * 96 ** 3 and e["d2h_bytes_per_step"] == 8 * 96 ** 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] > d["value"]
    assert d["clocks"]["samples"] >= 0 and "reasons" in d["clocks"]
    assert d["uncompressed"]["iterations"] == it
    assert d["spmv_7pt_256"]["frac_of_nominal_8TBs"] > 0.75             # the SpMV target of the metric
