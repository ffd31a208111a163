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
           "--steps", "2", "--warmup", "3", "--cpu-planes", "16", "--cpu-its", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
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
    pa = d["parity"]
    assert pa["iterations"] == it and pa["meets_bar"] is True and pa["status"] == 0
    assert pa["true_relres"] <= 1e-10 and pa["x_norm2"] > 0 and len(pa["x_at_rows"]) == 8
    assert d["e2e"]["true_relres"] <= 1e-10 and d["e2e"]["iterations"] == it
    assert d["gpu_launches"] >= 3 * 2 * it                              # 3 kernels per iteration, 2 steps
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000
    assert rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["algorithmic_bytes_per_launch"] == 12 * d["config"]["nnz"] + 4 * (96 ** 3 + 1) + 16 * 96 ** 3
    # frac is the DRAM-side figure (stored bytes); the algorithmic one sits beside it
    assert rf["bytes_per_launch"] < rf["algorithmic_bytes_per_launch"]   # index compression is on
    assert rf["frac"] < rf["frac_algorithmic"]
    e = d["e2e"]
    assert e["unit"] == "s" and e["value"] >= d["value"] * 0.9
    assert e["h2d_bytes_per_step"] == 2 * 8 * 96 ** 3 and e["d2h_bytes_per_step"] == 8 * 96 ** 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] > d["value"]
    assert d["clocks"]["samples"] >= 0 and "reasons" in d["clocks"]
    assert d["uncompressed"]["iterations"] == it
    assert d["spmv_7pt_256"]["frac_of_nominal_8TBs"] > 0.75             # the SpMV target of the metric
    # every BASELINE.json config in the driver-run line
    p7 = d["pcg_7pt_256"]
    assert p7["status"] == 0 and p7["true_relres"] <= 1e-10 and 900 < p7["iterations"] < 1100
    assert set(d["nek"]) == {"tj7a_A_12", "tj7a_A_15", "tj7a_A_18", "xn3b_A_10", "xn3b_A_12", "xn3b_A_15",
                             "xn3b_A_18"}
    for name, o in d["nek"].items():
        for leg in ("onchip", "streaming"):
            assert o[leg]["status"] == 0 and o[leg]["true_relres"] <= 1e-10 and o[leg]["rel_diff_direct"] <= 1e-8
        assert o["onchip"]["path"] == 1 and o["streaming"]["path"] == 0
    pl = d["powerlaw_50m"]
    assert pl["auto"]["ms_per_spmv"] > 0 and pl["column_blocked"]["col_blocks"] > 1
