"""bench.py's contract, the parts that can be pinned without a GPU: the metric and workload text both arms print, the
defaults the driver's bare command runs with, the flop conventions the roofline is computed from,
and the reference arm end to end on a tiny budget (it is CPU-only by construction)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from pimc_b200 import synth  # noqa: E402


def test_metric_and_workload_are_baselines():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith("ISF+S(q) evaluations/sec (N=256 He-4")
    assert bench.METRIC.startswith("ISF+S(q) evaluations/sec (N=256 He-4") and bench.UNIT == "evaluations/s"
    s = synth.SHAPES["C2"]
    assert (s.N, s.M, s.ndim, s.nq) == (256, 170, 3, 64)
    assert bench.metric_name(s, 64) == bench.METRIC
    cfg = bench.shared_config(s, 64)
    assert set(cfg) == {"workload"} and "N=256 M=170 nq=64 ndim=3" in cfg["workload"]      # no model keys; identical in both arms


def test_defaults_of_the_bare_command(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert a.gpus == 1 and a.impl == "ours" and a.workload == "C2"
    assert a.steps >= 20 and a.warmup >= 3                       # timing rules: W >= 3
    assert a.batches_per_step >= 16                              # one step = one output bin of >= 16 batches
    assert a.collective == "lib" and a.exchange == "pipelined" and a.shard == "config"


def test_flop_conventions_match_the_survey():
    s = synth.SHAPES["C2"]
    rho, corr = bench.algorithmic_flops(s, 64)
    assert rho == 64 * 256 * 170 * (2 * 3 + 40 + 2) and corr == 64 * 170 * (170 // 2 + 1) * 4
    assert abs(rho - 1.337e8) < 1e5 and abs(corr - 3.74e6) < 1e4                          # 1.337e8 + 3.74e6
    # the generic kernel is accounted in the survey's convention, the lattice kernels in useful flop of what they execute
    assert bench.kernel_algorithmic_flops(s, 64, {"path": 0}) == rho
    plan = {"path": 1, "nmax_x": 2, "nmax_y": 2, "nmax_z": 2, "L_rows": 22, "groups": 19}
    per_bead = 3 * 41 + 6 * 3 + 8 * ((22 - 1) // 4) + 2 * 8 * 19
    assert bench.kernel_algorithmic_flops(s, 64, plan) == 256 * 170 * per_bead
    assert bench.kernel_algorithmic_flops(s, 64, plan) < rho                               # the factorised form does less work


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` on a tiny CPU budget: one JSON line with the contract's keys, the same `config` the GPU
    arm would print, zero launches and zero copied bytes."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1.0"], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["dtype"] == "f64"
    assert d["config"] == bench.shared_config(synth.SHAPES["C2"], 64)
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["extrapolated"] is True
