"""One short GPU pass over the on-demand entry points added after the step path (sph_conserved,
sph_column_density, the 100k Sod tube): each check prints one JSON line and appends it to gpurun_out/quickcheck.jsonl
as soon as it is done.  No torch import (the engine is ctypes + CUDA only), so it fits in a minute of box time:

    python tests/gpu_quickcheck.py [conserved] [image] [merger] [sod_variable] [sod_fixed] [ring]
"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out", "quickcheck.jsonl")


def emit(name, t0, **kw):
    rec = dict(check=name, seconds=round(time.perf_counter() - t0, 3), **kw)
    line = json.dumps(rec, default=float)
    print(line, flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def check_conserved():
    from summersph_b200 import default_params, MODE_VARIABLE_H, ics
    from summersph_b200.engine import Engine
    from oracle.oracle import Oracle
    from test_widen_conserved import assert_same_sums
    t0 = time.perf_counter()
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(4000, seed=42)
    o = Oracle(p); o.upload(b, s)
    with Engine(p) as e:
        e.upload(b, s)
        g0, o0 = e.conserved(), o.conserved()
        assert_same_sums(g0, o0)
        dto, to = o.step(0.01, 0.0); dte, te = e.step(0.01, 0.0)
        assert (dto, to) == (dte, te)
        g1, o1 = e.conserved(), o.conserved()
        assert_same_sums(g1, o1)
    emit("conserved", t0, ok=True, engine_first=g0, oracle_first=o0, engine_after_step=g1, oracle_after_step=o1)


def check_image():
    from summersph_b200 import default_params, MODE_VARIABLE_H, ics
    from summersph_b200.engine import Engine
    from test_widen_density_image import image_ref
    t0 = time.perf_counter()
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(800, seed=12)
    extent, shape = (-110.0, 90.0, -60.0, 120.0), (48, 40)
    with Engine(p) as e:
        e.upload(b, s)
        img = e.column_density("z", extent, shape)
        imx = e.column_density("x", extent, shape)
    ref = image_ref(b.x, b.y, b.m, b.h, extent, shape)
    rex = image_ref(b.y, b.z, b.m, b.h, extent, shape)
    err = float(np.max(np.abs(img - ref)) / ref.max()); erx = float(np.max(np.abs(imx - rex)) / rex.max())
    emit("image", t0, ok=bool(err < 1e-9 and erx < 1e-9), rel_err_z=err, rel_err_x=erx, mass=float(img.sum() * (200 / 40) * (180 / 48)), mass_true=float(b.m.sum()))


def check_sod(kind):
    from summersph_b200.engine import Engine
    from sod_report import run_sod, sod_case
    t0 = time.perf_counter()
    p, b, s, geom = sod_case(kind, 100_000)
    with Engine(p) as e:
        rep = run_sod(e, b, s, geom)
    emit("sod_100k_" + kind, t0, ok=True, **rep)


def check_merger():
    from summersph_b200 import default_params, MODE_VARIABLE_H, FLAG_SINK_MERGE_SPIN, ics, Sinks
    from summersph_b200.engine import Engine
    from oracle.oracle import Oracle
    from test_widen_sink_merger import compare_sinks, orbital_L, sinks_L
    t0 = time.perf_counter()
    p = default_params(MODE_VARIABLE_H | FLAG_SINK_MERGE_SPIN, bounding_size=85.0)
    b, _ = ics.keplerian_disc(8_000, seed=9)
    s = Sinks([0.0, 40.0, 43.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.2], [0.0, 6.0, 5.5], [0.0, 0.0, 0.0],
              [1.0, 0.01, 0.02], [13.0, 6.0, 2.0])
    L_scale = np.linalg.norm(orbital_L(b.m, b.x, b.y, b.z, b.vx, b.vy, b.vz)) + np.linalg.norm(sinks_L(s))
    o = Oracle(p); o.upload(b, s)
    with Engine(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert o.sizes() == e.sizes() and (dto, to) == (dte, te)
            compare_sinks(e, o, L_scale)
        ce, co = e.conserved(), o.conserved()
        emit("merger", t0, ok=True, sizes=list(e.sizes()), spin_engine=e.sink_spin().tolist(), spin_oracle=o.sink_spin().tolist(),
             lz_engine=ce["lz"], lz_oracle=co["lz"])


def check_ring():
    """The 1M thin ring for 20 steps (tests/test_widen_thin_ring.py::test_ring_1M_20_steps) with its numbers printed."""
    import test_widen_thin_ring as TR
    from summersph_b200 import default_params, MODE_VARIABLE_H, ics
    from summersph_b200._abi import drift_report
    from summersph_b200.engine import Engine
    t0 = time.perf_counter()
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.thin_ring(TR.N_FULL)
    m0 = TR.ring_moments(b)
    with Engine(p) as e:
        e.upload(b, s)
        first = e.conserved()
        dt, t = 0.01, 0.0
        ts = time.perf_counter()
        for _ in range(TR.STEPS):
            dt, t = e.step(dt, t)
        wall = time.perf_counter() - ts
        last = e.conserved()
        b1, _ = e.download()
    emit("ring_1M", t0, ok=True, t=t, dt=dt, wall_s=wall, moments_first=m0, moments_last=TR.ring_moments(b1), drift=drift_report(first, last))


CHECKS = {"conserved": check_conserved, "ring": check_ring, "image": check_image, "merger": check_merger,
          "sod_variable": lambda: check_sod("variable"), "sod_fixed": lambda: check_sod("fixed")}

if __name__ == "__main__":
    which = sys.argv[1:] or list(CHECKS)
    rc = 0
    for name in which:
        try:
            CHECKS[name]()
        except Exception:
            rc = 1
            emit(name, time.perf_counter(), ok=False, error=traceback.format_exc()[-1500:])
    sys.exit(rc)
