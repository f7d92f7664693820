"""Sod shock tube against the exact Riemann solution (BASELINE.json configs[1]) — shared by the GPU test
(tests/test_sod_analytic.py: the CUDA engine at the full 100k particles) and by the command line below, which can
also run the CPU oracle so that the two L1 errors sit side by side ("compared against the analytic solution and
the reference").  Lives under tests/ because it may load the oracle; the product never imports it.

    python tests/sod_report.py [--n 100000] [--t-end 0.2] [--engine | --oracle [--threads 8]] [--out report.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from summersph_b200 import default_params, MODE_VARIABLE_H, EVAL_TREE, EVAL_DENSITY, ics   # noqa: E402
from summersph_b200._abi import drift_report                                               # noqa: E402
from summersph_b200.analytic import sod_core_mask, sod_l1_errors                           # noqa: E402

RHO_SCALE = 1e-9


def run_sod(sim, bodies, sinks, geom, gamma=1.4):
    """Advance `sim` (Engine or Oracle: same interface) with the reference's loop (dt0 = 1e-2, no final-step
    clipping, SUMMER_SPH.f90:872-879) to the first t >= geom['t_end'], then compare the core particles with the
    exact solution at that t.  Returns the report dict."""
    sim.upload(bodies, sinks)
    first = sim.conserved()
    dt, t, steps = 1.0e-2, 0.0, 0
    t0 = time.perf_counter()
    while t < geom["t_end"]:
        dt, t = sim.step(dt, t)
        steps += 1
    wall = time.perf_counter() - t0
    last = sim.conserved()
    sim.evaluate(EVAL_TREE | EVAL_DENSITY)             # rho and P of the final state (F:894-896)
    b, _ = sim.download()
    d = sim.diag()
    mask = sod_core_mask(b.x, b.y, b.z, t, geom, gamma)
    rep = sod_l1_errors(b.x, d["rho"], b.vx, d["P"], t, mask, gamma, RHO_SCALE)
    # plateau values between contact and shock / contact and rarefaction tail (exact: 0.26557 | 0.42632, v 0.92745, P 0.30313)
    xs = b.x[mask] / t
    post = mask.copy(); post[mask] = (xs > 1.15) & (xs < 1.5)
    star_l = mask.copy(); star_l[mask] = (xs > 0.15) & (xs < 0.7)
    rep.update(
        n=len(b), steps=steps, t=t, dt_last=dt, wall_s=wall,
        rho_post_shock=float(np.mean(d["rho"][post]) / RHO_SCALE) if post.any() else None,
        rho_star_left=float(np.mean(d["rho"][star_l]) / RHO_SCALE) if star_l.any() else None,
        v_star=float(np.mean(b.vx[post | star_l])) if (post | star_l).any() else None,
        p_star=float(np.mean(d["P"][post | star_l]) / RHO_SCALE) if (post | star_l).any() else None,
        drift={k: v for k, v in drift_report(first, last).items()},
        geom={k: float(v) for k, v in geom.items()},
    )
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--t-end", type=float, default=0.2)
    ap.add_argument("--oracle", action="store_true", help="run the CPU oracle instead of the CUDA engine")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    p = default_params(MODE_VARIABLE_H)
    b, s, geom = ics.sod_box(a.n, a.t_end, rho_scale=RHO_SCALE)
    if a.oracle:
        from oracle.oracle import Oracle
        sim = Oracle(p, threads=a.threads)
        impl = f"oracle ({a.threads} threads: OpenMP sums differ from the serial order at rounding level)"
    else:
        from summersph_b200.engine import Engine
        sim = Engine(p)
        impl = "engine (CUDA, C-ABI)"
    rep = run_sod(sim, b, s, geom)
    rep["impl"] = impl
    sim.close()
    line = json.dumps(rep)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
