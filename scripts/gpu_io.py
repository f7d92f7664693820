"""Developer script: time sph_upload / sph_download with pinned vs pageable host buffers."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from summersph_b200 import default_params, MODE_VARIABLE_H, ics, Bodies, Sinks
from summersph_b200.state import GAS_FIELDS
from summersph_b200.engine import Engine
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16_000_000
p = default_params(MODE_VARIABLE_H)
b, s = ics.keplerian_disc(n)
pin = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in GAS_FIELDS}
for k in GAS_FIELDS: pin[k].numpy()[:] = getattr(b, k)
hb = Bodies(*[pin[k].numpy() for k in GAS_FIELDS])
opin = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in GAS_FIELDS}
ob = Bodies(*[opin[k].numpy() for k in GAS_FIELDS])
e = Engine(p)
for name, src in (("pageable", b), ("pinned", hb)):
    for rep in range(3):
        t0 = time.perf_counter(); e.upload(src, s); t1 = time.perf_counter()
        print(f"upload {name}: {1e3*(t1-t0):.1f} ms")
e.evaluate()
for rep in range(3):
    t0 = time.perf_counter(); e.download(into=(ob, Sinks.empty(1))); t1 = time.perf_counter()
    print(f"download pinned: {1e3*(t1-t0):.1f} ms")
t0 = time.perf_counter(); e.download(); t1 = time.perf_counter(); print(f"download pageable(new arrays): {1e3*(t1-t0):.1f} ms")
