"""Developer script: engine vs oracle on a small disc, prints max errors per quantity (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, EVAL_ALL, EVAL_TREE, EVAL_DENSITY, EVAL_GRAVITY, EVAL_SINKS, EVAL_SPH
from summersph_b200 import ics
from summersph_b200.engine import Engine
from oracle.oracle import Oracle


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    scale = np.maximum(np.abs(b), np.sqrt(np.mean(b * b)) + 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def run(mode, n, steps=3):
    print(f"=== mode={mode} n={n}")
    p = default_params(mode)
    b, s = ics.keplerian_disc(n)
    o = Oracle(p); o.record_neighbours(True); o.upload(b, s)
    e = Engine(p); e.upload(b, s)
    t0 = time.time(); o.evaluate(); t1 = time.time(); e.evaluate(); t2 = time.time()
    print(f"oracle eval {t1-t0:.3f}s engine eval {t2-t1:.3f}s stages {e.stage_times()}")
    to, te = o.tree(), e.tree()
    print("order equal:", np.array_equal(to["order"], te["order"]), "level equal:", np.array_equal(to["level"], te["level"]))
    for k in ("cx", "cy", "cz", "size"):
        print(f"  leaf {k} bit-equal:", np.array_equal(to[k], te[k]))
    co, ho, oo, lo = o.neighbours()
    ce, he, oe, _ = e.neighbours(with_list=False)
    ce2, he2, oe2, _ = e.neighbours(with_list=False)
    print("ngb count equal:", np.array_equal(co, ce), "hash equal:", np.array_equal(ho, he), "total", int(oo[-1]), int(oe[-1]), "repeatable:", np.array_equal(ce, ce2), int(oe2[-1]))
    bad = np.nonzero(co != ce)[0]
    print("  mismatching rows:", bad.size, bad[:10], co[bad[:10]], ce[bad[:10]])
    try:
        ce, he, oe, le = e.neighbours()
        print("  list equal:", np.array_equal(lo, le))
    except Exception as ex:
        print("  list failed:", ex)
    do, de = o.diag(), e.diag()
    for k in do:
        print(f"  {k:10s} relerr {relerr(de[k], do[k]):.3e}")
    print("counters oracle", o.counters()); print("counters engine", e.counters())
    # phase isolation
    for name, mask in (("grav", EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY), ("sinks", EVAL_TREE | EVAL_DENSITY | EVAL_SINKS), ("sph", EVAL_TREE | EVAL_DENSITY | EVAL_SPH)):
        o.evaluate(mask); e.evaluate(mask)
        do, de = o.diag(), e.diag()
        print(f"  [{name}] " + " ".join(f"{k}={relerr(de[k], do[k]):.2e}" for k in ("ax", "ay", "az", "udot", "alphadot", "sink_ax")))
    o.record_neighbours(False)
    o.upload(b, s); e.upload(b, s)
    dto = dte = 0.01; tto = tte = 0.0
    for k in range(steps):
        dto, tto = o.step(dto, tto); dte, tte = e.step(dte, tte)
        bo, so = o.download(); be, se = e.download()
        print(f"step {k}: dt {dto} {dte} t {tto} {tte} n {o.sizes()} {e.sizes()} hiter {o.counters()['h_iterations']} {e.counters()['h_iterations']}")
        if len(bo) == len(be):
            print("   " + " ".join(f"{f}={relerr(getattr(be, f), getattr(bo, f)):.2e}" for f in ("x", "y", "z", "vx", "vy", "vz", "u", "alpha", "h")))
            print("   sinks " + " ".join(f"{f}={relerr(getattr(se, f), getattr(so, f)):.2e}" for f in ("x", "vx", "m")))
    print("stage ms", e.stage_times(), "launches", e.launch_count())


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    run(MODE_FIXED_H, n)
    run(MODE_VARIABLE_H, n)
