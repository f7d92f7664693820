"""Developer script: per-stage times of N-particle device-generated disc steps; with a GW_STATS variant build the
gravity walk's own counters appear on stderr.  usage: gpu_stats.py [N] [steps] [variant-name]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summersph_b200 import default_params, MODE_VARIABLE_H
from summersph_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = None
if len(sys.argv) > 3:
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "summersph_b200", "variants", f"libsph_{sys.argv[3]}.so")
e = Engine(default_params(MODE_VARIABLE_H), lib_path=lib)
e.ics_disc(n)
dt, t = 0.01, 0.0
for k in range(steps):
    e.timer_start()
    dt, t = e.step(dt, t)
    ms = e.timer_stop()
    print(f"step {k}: {ms:.1f} ms dt={dt} n={e.sizes()} far_reuse={e.far_reuse_count()}")
    print("   stages:", {k_: round(v, 2) for k_, v in e.stage_times().items()}, flush=True)
