"""Developer script: the host-buffer step (sph_step_host) against upload + step + download, wall clock, pinned buffers.
usage: gpu_io.py [N] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summersph_b200 import default_params, MODE_VARIABLE_H, Bodies, Sinks
from summersph_b200.state import GAS_FIELDS
from summersph_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = Engine(default_params(MODE_VARIABLE_H))
e.ics_disc(n)
pin = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in GAS_FIELDS}
b = Bodies(*[pin[k].numpy() for k in GAS_FIELDS]); s = Sinks.empty(1)
e.download(into=(b, s))
dt, t = e.step(0.01, 0.0)
e.download(into=(b, s))
for name, env in (("separate", None), ("fused", {}), ("fused, no late columns", {"SPH_B200_IO_NO_LATE": "1"}), ("fused, no early columns", {"SPH_B200_IO_NO_EARLY": "1"})):
    for k in ("SPH_B200_IO_NO_LATE", "SPH_B200_IO_NO_EARLY"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    os.environ["SPH_B200_IO_TRACE"] = "1"
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        if env is None:
            e.upload(b, s); t1 = time.perf_counter(); dt, t = e.step(dt, t); t2 = time.perf_counter(); e.download(into=(b, s))
            print(f"   upload {1e3*(t1-t0):.1f} step {1e3*(t2-t1):.1f} download {1e3*(time.perf_counter()-t2):.1f}")
        else:
            so = Sinks.empty(9)
            dt, t, n2, ns2 = e.step_host(b, s, dt, t, into=(b, so))
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{name}: {['%.1f' % v for v in ts]} ms per step", flush=True)
    print("   stages:", {k_: round(v, 1) for k_, v in e.stage_times().items() if v > 0.3})
