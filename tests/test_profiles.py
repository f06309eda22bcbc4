"""The committed measurement artefacts under profiles/ are well-formed: every bench line parses as ONE JSON object carrying the
keys the bench contract names, and the per-kernel ncu summaries carry durations and stall breakdowns.  (CPU-only; the numbers
themselves are produced on the B200 by bench.py / tools/ncu_summary.py.)"""
import glob
import json
import os

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")
R02_LINES = sorted(glob.glob(os.path.join(PROFILES, "r02_bench_*.json")))


def _last_line(path):
    with open(path) as f:
        lines = [ln for ln in f.read().strip().splitlines() if ln.strip()]
    return json.loads(lines[-1])


def test_round2_bench_lines_exist():
    names = {os.path.basename(p) for p in R02_LINES}
    for need in ("r02_bench_c2.json", "r02_bench_c2_n2.json", "r02_bench_c2_n4.json", "r02_bench_c2_n8.json", "r02_bench_c2f32.json",
                 "r02_bench_c3.json", "r02_bench_c4.json", "r02_bench_c5_r8.json", "r02_bench_c5_r64_n8.json",
                 "r02_bench_reference.json", "r02_bench_hmm_h1.json", "r02_bench_hmm_h2.json", "r02_bench_hmm_h3.json"):
        assert need in names, need


@pytest.mark.parametrize("path", R02_LINES, ids=[os.path.basename(p) for p in R02_LINES])
def test_bench_line_contract(path):
    d = _last_line(path)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config"):
        assert key in d, key
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
        return
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] in ("hbm", "tensor") and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if d["n_gpus"] > 1 and "parity_multi" in d:
        pm = d["parity_multi"]
        assert pm["ranks_bit_identical"] is True and pm["max_rel"] <= 1e-9


def test_ncu_summaries_are_complete():
    files = sorted(glob.glob(os.path.join(PROFILES, "r02_*_ncu_summary.json")))
    assert len(files) >= 5
    for path in files:
        d = json.load(open(path))
        assert d["kernels"], path
        for k in d["kernels"]:
            assert k["duration_ns"] > 0 and k["stall_breakdown_pct"] and k["sass_hot_spots"], (path, k["kernel"])


def test_traffic_file_matches_the_bench_lookup():
    t = json.load(open(os.path.join(PROFILES, "traffic.json")))
    for key in ("c2_n1", "c3_n1", "c2f32_n1"):
        assert t[key] > 1e9
    assert 0 < t["tensor_pipe_active_pct"]["c2_n1"] <= 100
