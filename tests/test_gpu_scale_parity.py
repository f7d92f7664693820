"""Parity at the size the metric is quoted on (BASELINE configs[3]: 16M-particle Keplerian disc + sink, variable h).

The full serial pair loop of the oracle would take hours at 16M, so the engine is held to the oracle three ways:
  * the tree of the whole set: Morton (depth-first leaf) order and every particle's leaf level / cell - bit-exact;
  * a random sample of targets evaluated by the oracle on the full 16M tree (oracle `sample_eval`: density, Omega, EOS,
    Barnes-Hut gravity, sink gravity, gather-form pair sums, alpha rate): rho Omega P c a du/dt dalpha/dt within 1e-10
    relative (scale max(|value|, sample RMS)), neighbour counts and hashes bit-exact;
  * the same after one full loop body (positions moved, h iterated, second tree).
Reference semantics: "SUMMER_SPH - Variable.f90":1120-1162.  SPH_SCALE_PARITY_N overrides the size (developer runs)."""
import os
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, ics

pytestmark = pytest.mark.gpu
N = int(float(os.environ.get("SPH_SCALE_PARITY_N", "16e6")))
N_SAMPLE = 2000
TOL = 1e-10


def _rel(a, b):
    scale = np.maximum(np.abs(b), np.sqrt(np.mean(b * b)))
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale))


def _check(e, o, b, s, tag):
    """engine state == (b, s); compare one evaluation of it with the oracle's sampled evaluation"""
    from oracle.oracle import Oracle  # noqa: F401
    n = len(b)
    o.upload(b, s)
    rng = np.random.default_rng(20251018)
    targets = np.sort(rng.choice(n, N_SAMPLE, replace=False)).astype(np.int32)
    so = o.sample_eval(targets)
    e.set_exact_counters(True)
    e.evaluate()
    e.set_exact_counters(False)
    # whole-set tree: order, levels, cells
    to, te = o.tree(), e.tree()
    assert np.array_equal(to["order"], te["order"]), f"{tag}: Morton order differs"
    assert np.array_equal(to["level"], te["level"]), f"{tag}: leaf levels differ"
    for k in ("cx", "cy", "cz", "size"):
        assert np.array_equal(to[k], te[k]), f"{tag}: leaf cell {k} differs"
    del to, te
    ce, he, _, _ = e.neighbours(with_list=False)
    assert np.array_equal(ce[targets], so["count"]), f"{tag}: neighbour counts differ"
    assert np.array_equal(he[targets], so["hash"]), f"{tag}: neighbour sets differ"
    d = e.diag()
    worst = {k: _rel(d[k][targets], so[k]) for k in ("rho", "omega", "P", "c", "ax", "ay", "az", "udot", "alphadot")}
    assert max(worst.values()) < TOL, (tag, worst)
    return worst, int(so["pairs"].sum())


def test_sampled_parity_at_benchmark_scale(built_engine):
    from summersph_b200.engine import Engine
    from oracle.oracle import Oracle
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(N, seed=20251018)
    s.radius[:] = p.sink_radius
    b.alpha[:] = 0.5                                  # viscosity terms on
    o = Oracle(p, threads=os.cpu_count() or 1)
    report = {}
    with Engine(p, device=0) as e:
        e.upload(b, s)
        report["initial"] = _check(e, o, b, s, "initial state")
        # one full loop body on the engine, then the same check on the state it produced
        e.upload(b, s)
        dt, t = e.step(0.01, 0.0)
        b1, s1 = e.download()
        assert len(b1) == e.sizes()[0]
        report["after_step"] = _check(e, o, b1, s1, "after one step")
    print("sampled parity at N =", N, report)
