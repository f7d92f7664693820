"""Developer script: per-stage timing of the engine at size N (V mode disc)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H
from summersph_b200 import ics
from summersph_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = MODE_VARIABLE_H if (len(sys.argv) <= 3 or sys.argv[3] == "V") else MODE_FIXED_H
p = default_params(mode)
t0 = time.time(); b, s = ics.keplerian_disc(n); print(f"ICs {time.time()-t0:.1f}s", flush=True)
if mode == MODE_FIXED_H:
    p.h_fixed = float(np.median(b.h))
e = Engine(p)
print("fp64 peak TFLOP/s", e.fp64_peak())
t0 = time.time(); e.upload(b, s); print(f"upload {time.time()-t0:.2f}s", flush=True)
dt, t = 0.01, 0.0
for k in range(steps):
    e.timer_start(); t0 = time.time()
    dt, t = e.step(dt, t)
    ms = e.timer_stop()
    st = e.stage_times(); c = e.counters()
    print(f"step {k}: {ms:.1f} ms (wall {1e3*(time.time()-t0):.1f}) dt={dt} n={e.sizes()} -> {n/ms*1e3:.3e} particle-steps/s")
    print("   stages:", {k_: round(v, 2) for k_, v in st.items()})
    print("   counters:", c, "groups", e.group_count(), flush=True)
t0 = time.time(); e.download(); print(f"download {time.time()-t0:.2f}s")
