"""Shared helpers for the golden-fixture and parity tests."""
import glob
import os
import numpy as np

from summersph_b200 import SphParams, Bodies, Sinks
from summersph_b200.state import GAS_FIELDS, SINK_FIELDS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    p = SphParams.from_buffer_copy(z["params"].tobytes())
    b = Bodies(*[z["in_" + k] for k in GAS_FIELDS])
    s = Sinks(*[z["in_sink_" + k] for k in SINK_FIELDS])
    return z, p, b, s


def run_case(sim, z, b, s):
    """Evaluate + step `sim` (Oracle or Engine: same interface) like make_golden.py; returns dict of outputs."""
    out = {}
    sim.upload(b, s)
    sim.evaluate()
    for k, v in sim.diag().items():
        out["ev_" + k] = v
    t = sim.tree()
    for k in ("order", "level", "cx", "cy", "cz", "size"):
        out["tree_" + k] = t[k]
    out["ngb_count"], out["ngb_hash"], _, _ = sim.neighbours(with_list=False)
    sim.upload(b, s)
    dt, t_ = 0.01, 0.0
    for _ in range(int(z["steps"][0])):
        dt, t_ = sim.step(dt, t_)
    b2, s2 = sim.download()
    for k in GAS_FIELDS:
        out["st_" + k] = getattr(b2, k)
    for k in SINK_FIELDS:
        out["st_sink_" + k] = getattr(s2, k)
    out["st_dt_t"] = np.array([dt, t_])
    return out
