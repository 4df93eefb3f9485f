"""Host-side checks of bench.py (no GPU): the JSON contract pieces that do not need a device."""
import json
import os
import subprocess
import sys
import time

import bench
from util import ROOT


def test_workload_config_names_the_workload():
    c = bench.workload_config(4)
    assert "orszag_tang 4096x4096 cells per GPU" in c["workload"] and c["ny_global"] == 4 * 4096
    assert c["partition"] == "y-slabs x4" and c["es_roundtrip"] == 1
    assert "model" not in c


def test_measured_peak_source():
    peak, how = bench.measured_peaks()
    assert peak > 1000 and how in ("measured", "fallback")
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        assert how == "measured"


def test_clock_sampler_parses_nvidia_smi_lines():
    s = bench.ClockSampler(0)
    t = time.time()
    s.samples = [(t, "1965, 1965, 431.20, Not Active, Not Active, Not Active, Active"),
                 (t + 0.1, "1950, 1965, 470.00, Not Active, Not Active, Not Active, Not Active"),
                 (t + 0.2, "[N/A], x"), (t + 100, "1000, 1965, 100.0, Active, Not Active, Not Active, Not Active")]
    out = s.summary(t - 1, t + 1)
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 2
    assert bench.ClockSampler(0).summary(0, 1) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_reference_arm_other_ranks_print_nothing():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_baseline_object_shape():
    b = bench.cpu_baseline(budget_s=1.0)
    assert b["kind"] == "port" and b["unit"] == "cell-updates/s" and b["cores"] >= 1 and b["value"] > 0
    assert "orszag_tang" in b["sample"] and json.dumps(b)


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every host core, report how
    many, and time the same sample at every N (round-1 SCALE record: 1 thread at N >= 2)."""
    lines = {}
    for n in (1, 4):
        env = dict(os.environ, RANK="0", WORLD_SIZE=str(n), OMP_NUM_THREADS="1", WENO5_REF_SAMPLE_N="128")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(n),
                            "--steps", "1"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        lines[n] = json.loads(r.stdout.strip().splitlines()[-1])
    for n, d in lines.items():
        assert d["impl"] == "reference" and d["n_gpus"] == n
        assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1), d["cpu_baseline"]
        assert "128x128" in d["cpu_baseline"]["sample"] and "warning" not in d
    assert lines[1]["cpu_baseline"]["sample"] == lines[4]["cpu_baseline"]["sample"]
