"""Conservation over whole runs of BASELINE config[0] (Keplerian disc, 10 000 gas + central sink) ON THE GPU ENGINE.

  python scripts/drift_report.py OUT.md [n]

`simulate(..., drift={})`: the reference's loop shell, conserved sums (`sph_conserved`) before the first and after the last step.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from summersph_b200 import default_params, ics, MODE_VARIABLE_H, MODE_FIXED_H   # noqa: E402
from summersph_b200.simulate import simulate                                    # noqa: E402


def main():
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    rows = []
    for mode, name in ((MODE_VARIABLE_H, "variable_h"), (MODE_FIXED_H, "fixed_h")):
        for end in (0.1, 1.0, 20.0):
            p = default_params(mode, end_time=end)
            b, s = ics.keplerian_disc(n, seed=20251018)
            s.radius[:] = p.sink_radius
            d = {}
            bb, ss, t, dt, steps = simulate(b, s, p, drift=d, log=lambda *_: None)
            f = lambda v: "n/a" if v is None else f"{v:+.2e}"   # noqa: E731
            rows.append(f"| {name} | {end} | {steps} | {t:.4f} | {len(bb.x)} | {f(d['energy_rel'])} | {f(d['momentum_rel'])} | {f(d['angular_momentum_rel'])} | {f(d['mass_rel'])} |")
            print(rows[-1], flush=True)
    with open(out, "w") as fh:
        fh.write(f"# Conservation over whole runs of BASELINE config[0] (Keplerian disc, {n} gas + central sink) - CUDA engine on the B200\n\n"
                 "`python scripts/drift_report.py` = `simulate(..., drift={})`: the reference's loop shell around `sph_step`, conserved sums from `sph_conserved`\n"
                 "before the first and after the last step.  dE relative to |E_total(0)|, dP relative to sqrt(2 E_kin M), dL relative to |L|.  The disc is cold and\n"
                 "light (M_disc = 0.01 M_sun), so E_total is dominated by the sink's potential; Barnes-Hut monopole gravity is not momentum conserving.\n\n"
                 "| program | end_time | steps | t reached | gas left | dE/\\|E0\\| | \\|dP\\|/sqrt(2 E_kin M) | \\|dL\\|/\\|L\\| | dM/M0 |\n|---|---|---|---|---|---|---|---|---|\n")
        fh.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
