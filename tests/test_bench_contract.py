"""bench.py's reference arm (the CPU port of the path on the host cores) prints the contract's JSON line.
CPU only: the b200 arm needs a GPU and is run by the driver / tests -m gpu environment."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-particles", "3000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-steps/s" and line["unit"] == "particle-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--particles", "1000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_recorded_b200_lines_carry_the_contract():
    """The B200 arm cannot run here; the lines it printed on the hardware (profiles/r2b_bench_*.json, copied unedited from the
    runs) must carry every key the driver's contract names, with consistent numbers."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2b_bench_16M_*gpu*.json")))
    assert files
    for f in files:
        line = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in line, (f, k)
        assert line["metric"] == "particle-steps/s" and line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
        assert "workload" in line["config"] and line["gpu_launches"] > 0
        n = line["config"]["particles"]
        assert abs(line["value"] - n * 1e3 / line["ms_per_step"]) < 1e-6 * line["value"]          # whole-job throughput of the timed steps
        e2e = line["e2e"]
        assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] >= 80 * n and e2e["d2h_bytes_per_step"] >= 80 * n      # every row, ten FP64 columns, both ways
        assert e2e["value"] < line["value"] and e2e["cold"]["value"] < e2e["value"]
        r = line["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (f, k)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        c = line["clocks"]
        assert c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if line["n_gpus"] == 1 and "cpu_baseline" in line:
            cb = line["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
