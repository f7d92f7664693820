"""Generates tests/golden/*.npz from the CPU oracle (oracle/sph_oracle.cpp).

The reference ships no golden vectors and cannot be built here (no Fortran compiler), so these fixtures
pin the ORACLE's output (regression pin for the restatement, and the small-case target for the CUDA
engine), not the Fortran program's.  The second, literal restatement (oracle/pyref.py) reproduces every file bit for bit
(tests/test_oracle_pyref.py::test_pyref_reproduces_golden).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, ics, Sinks   # noqa: E402
from summersph_b200.state import GAS_FIELDS, SINK_FIELDS                                 # noqa: E402
from oracle.oracle import Oracle                                                        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, mode, bodies, sinks, steps=2, **pk):
    p = default_params(mode, **pk)
    o = Oracle(p)
    o.record_neighbours(True)
    o.upload(bodies, sinks)
    o.evaluate()
    out = {}
    for k in GAS_FIELDS:
        out["in_" + k] = getattr(bodies, k)
    for k in SINK_FIELDS:
        out["in_sink_" + k] = getattr(sinks, k)
    out["params"] = np.frombuffer(bytes(p), dtype=np.uint8)
    for k, v in o.diag().items():
        out["ev_" + k] = v
    t = o.tree()
    for k in ("order", "level", "cx", "cy", "cz", "size"):
        out["tree_" + k] = t[k]
    cnt, hsh, off, lst = o.neighbours()
    out["ngb_count"], out["ngb_hash"] = cnt, hsh
    c = o.counters()
    out["counters"] = np.array([c[k] for k in sorted(c)], np.int64)
    o.record_neighbours(False)
    o.upload(bodies, sinks)
    dt, t_ = 0.01, 0.0
    for _ in range(steps):
        dt, t_ = o.step(dt, t_)
    b2, s2 = o.download()
    for k in GAS_FIELDS:
        out["st_" + k] = getattr(b2, k)
    for k in SINK_FIELDS:
        out["st_sink_" + k] = getattr(s2, k)
    out["st_dt_t"] = np.array([dt, t_])
    out["steps"] = np.array([steps])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n", len(bodies), "->", len(b2), "dt", dt, "t", t_)


if __name__ == "__main__":
    b, s = ics.keplerian_disc(600, seed=101)
    s.radius[:] = 3.5
    case("disc600_fixed", MODE_FIXED_H, b, s)
    s5 = s.copy(); s5.radius[:] = 5.0
    case("disc600_variable", MODE_VARIABLE_H, b, s5)
    # accretion + bounds: big sink radius, tight bounding box
    s12 = s.copy(); s12.radius[:] = 14.0
    case("disc600_variable_accrete", MODE_VARIABLE_H, b, s12, steps=2, bounding_size=90.0)
    case("disc600_fixed_accrete", MODE_FIXED_H, b, s12, steps=2, bounding_size=90.0)
    bs, ss = ics.sod_tube(1500, width=3)
    case("sod_variable", MODE_VARIABLE_H, bs, ss, steps=2)
