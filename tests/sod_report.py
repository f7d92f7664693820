"""Sod shock tube against the exact Riemann solution (BASELINE.json configs[1]) — shared by the GPU test
(tests/test_widen_shock_tube.py: the CUDA engine at the full 100k particles) and by the command line below, which can
also run the CPU oracle so that the two L1 errors sit side by side ("compared against the analytic solution and
the reference").  Lives under tests/ because it may load the oracle; the product never imports it.

    python tests/sod_report.py [--n 100000] [--fixed] [--oracle [--threads 8]] [--out report.json]
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

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, EVAL_TREE, EVAL_DENSITY, ics   # noqa: E402
from summersph_b200._abi import drift_report                                               # noqa: E402
from summersph_b200.analytic import sod_core_mask, riemann_exact, riemann_star, effective_gas, OMEGA_LATTICE                         # noqa: E402

RHO_SCALE = 1e-9


def run_sod(sim, bodies, sinks, geom, gamma=1.4):
    """Advance `sim` (Engine or Oracle: same interface) with the reference's loop (dt0 = 1e-2, no final-step
    clipping, SUMMER_SPH.f90:872-879) to the first t >= geom['t_end'], then compare the core particles with the
    exact solution at that t.  Returns the report dict.

    Two exact solutions are reported.  `nominal`: Sod's problem for the ideal gas the ICs describe (gamma = 1.4).
    `effective`: the variable-h program divides every pressure term by its grad-h factor Omega (V:413-425), and its
    Omega (V:455,487: 1 + h/(3 rho) sum m (3W - r dW/dr)/h) is ~3 on a uniform lattice, not ~1 — SURVEY.md §8(a) #9
    keeps the sign as coded.  Momentum and energy equations both carry P/Omega, so the particles evolve an ideal
    gas with P_eff = P/Omega0 and gamma_eff = 1 + (gamma-1)/Omega0; Omega0 is measured on the initial lattice.
    In fixed-h mode there is no Omega (Omega0 = 1) and the two solutions coincide."""
    variable = bool(sim.params.mode & MODE_VARIABLE_H)
    sim.upload(bodies, sinks)
    omega0 = 1.0
    if variable:
        sim.evaluate(EVAL_TREE | EVAL_DENSITY)
        m0 = sod_core_mask(bodies.x, bodies.y, bodies.z, 0.0, geom, gamma)
        omega0 = float(np.median(sim.diag()["omega"][m0]))
        sim.upload(bodies, sinks)
    first = sim.conserved()
    dt, t, steps = 1.0e-2, 0.0, 0
    t0 = time.perf_counter()
    while t < geom["t_end"]:
        dt, t = sim.step(dt, t)
        steps += 1
    wall = time.perf_counter() - t0
    last = sim.conserved()
    sim.evaluate(EVAL_TREE | EVAL_DENSITY)             # rho of the final state (F:894-896)
    b, _ = sim.download()
    d = sim.diag()
    mask = sod_core_mask(b.x, b.y, b.z, t, geom, gamma)
    rho = d["rho"][mask] / RHO_SCALE; vx = b.vx[mask]; u = b.u[mask]; xi = b.x[mask] / t

    def errors(om):
        g_eff, st = effective_gas(gamma, om)
        re, ve, pe = riemann_exact(xi, gamma=g_eff, **st)
        ue = pe / ((g_eff - 1.0) * re)
        ps, vs = riemann_star(gamma=g_eff, **st)
        return {"gamma": g_eff, "p_star": ps * om, "v_star": vs,
                "rho_l1": float(np.mean(np.abs(rho - re)) / np.mean(re)),
                "v_l1": float(np.mean(np.abs(vx - ve)) / vs),
                "u_l1": float(np.mean(np.abs(u - ue)) / np.mean(ue))}

    rep = {
        "mode": "variable_h" if variable else "fixed_h", "n": len(b), "n_core": int(np.count_nonzero(mask)),
        "steps": steps, "t": t, "dt_last": dt, "wall_s": wall, "omega0": omega0,
        "nominal": errors(1.0),
        "effective": errors(omega0),
        "drift": drift_report(first, last),
        "geom": {k: float(v) for k, v in geom.items()},
    }
    return rep


def sod_case(kind, n=100_000):
    """The two canonical runs of config 2.  'variable': "SUMMER_SPH - Variable.f90", tube sized for the waves of its
    effective gas (Omega ~ 3) and run to t = 0.4 so that every h stays above the 0.01 below which that program
    never updates h (V:528).  'fixed': SUMMER_SPH.f90 with smoothing = eta * right-hand spacing, t = 0.2.
    Returns (params, bodies, sinks, geom)."""
    if kind == "variable":
        b, s, geom = ics.sod_box(n, 0.4, rho_scale=RHO_SCALE, omega=OMEGA_LATTICE)
        return default_params(MODE_VARIABLE_H), b, s, geom
    b, s, geom = ics.sod_box(n, 0.2, rho_scale=RHO_SCALE)
    return default_params(MODE_FIXED_H, h_fixed=geom["h_right"]), b, s, geom


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--oracle", action="store_true", help="run the CPU oracle instead of the CUDA engine")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--fixed", action="store_true", help="the fixed-h program instead of the variable-h one")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    p, b, s, geom = sod_case("fixed" if a.fixed else "variable", a.n)
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
