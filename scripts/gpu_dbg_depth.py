"""Developer script: NaN-pattern comparison on a depth-limited tree (multi-particle leaves)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from summersph_b200 import default_params, MODE_VARIABLE_H, ics
from summersph_b200.engine import Engine
from oracle.oracle import Oracle
p = default_params(MODE_VARIABLE_H, max_depth=4)
b, s = ics.keplerian_disc(6000, seed=4)
o = Oracle(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
for exact in (False, True):
    with Engine(p, exact_counters=exact) as e:
        e.upload(b, s); e.evaluate()
        do, de = o.diag(), e.diag()
        to = o.tree()
        for k in ("rho", "ax", "udot", "alphadot"):
            no, ne = np.isnan(do[k]), np.isnan(de[k])
            bad = np.nonzero(no != ne)[0]
            print(exact, k, "nan oracle", no.sum(), "engine", ne.sum(), "mismatch", len(bad))
            for i in bad[:8]:
                print("   i", i, "n_in_leaf", to["n_in_leaf"][i], "level", to["level"][i], "oracle", do[k][i], "engine", de[k][i], "rho_o", do["rho"][i], "rho_e", de["rho"][i])
        print(o.counters(), e.counters())
