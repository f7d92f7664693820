"""Randomised parity cases on the GPU: program, physical parameters (gamma, eta, h tolerance, timestep scale, bounding
cube), geometry (disc, ring, sphere with random velocities), viscosity and 1-3 massive sinks drawn per seed; two loop bodies
on the CUDA engine and on the oracle.  Benign regimes only (natural tree depth, every density positive): the NaN regimes of
shallow depth limits are covered by tests/test_gpu_parity.py's dedicated cases and, between the two CPU restatements, by
tests/test_oracle_pyref.py.  Bit-exact: dt, t, particle and sink counts, Morton order; 1e-10 relative: the state."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SOFT_USES_HI, ics, Sinks
from summersph_b200.state import GAS_FIELDS
from conftest import relerr

pytestmark = pytest.mark.gpu


def random_case(seed):
    rng = np.random.default_rng(7000 + seed)
    mode = [MODE_VARIABLE_H, MODE_FIXED_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI][seed % 3]
    p = default_params(mode, bounding_size=float(rng.choice([70.0, 90.0, 1500.0])), gamma=float(rng.choice([1.4, 5 / 3])),
                       eta=float(rng.uniform(1.1, 1.3)), convergence_criteria=float(rng.choice([1e-2, 1e-3, 1e-4])),
                       timestep_scale=float(rng.choice([0.1, 0.25, 0.5])))
    n = int(rng.integers(2000, 5000))
    kind = seed % 3
    if kind == 0:
        b, _ = ics.keplerian_disc(n, seed=seed)
    elif kind == 1:
        b, _ = ics.thin_ring(n, seed=seed)
    else:
        b, _ = ics.uniform_sphere(n, seed=seed, radius=60.0)
        b.vx[:] = rng.normal(0, 0.3, n); b.vy[:] = rng.normal(0, 0.3, n); b.vz[:] = rng.normal(0, 0.1, n)
    b.alpha[:] = rng.uniform(0.05, 1.0, n)
    if mode == MODE_FIXED_H:
        p = p.copy(h_fixed=float(np.median(b.h)) * 1.2)
    ns = int(rng.integers(1, 4))
    s = Sinks(rng.uniform(-30, 30, ns), rng.uniform(-30, 30, ns), rng.uniform(-1, 1, ns), rng.normal(0, 0.5, ns), rng.normal(0, 0.5, ns),
              rng.normal(0, .05, ns), rng.choice([0.01, 0.3, 1.0], ns), rng.uniform(2, 12, ns))
    return p, b, s


@pytest.mark.parametrize("seed", range(6))
def test_random_case_engine_follows_oracle(seed, built_engine):
    from summersph_b200.engine import Engine
    from oracle.oracle import Oracle
    p, b, s = random_case(seed)
    o = Oracle(p); o.upload(b, s)
    with Engine(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for k in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert (dto, to) == (dte, te), k
            assert o.sizes() == e.sizes(), k
        be, se = e.download(); bo, so = o.download()
        for k in GAS_FIELDS:
            assert relerr(getattr(be, k), getattr(bo, k)) < 1e-10, k
        for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
            assert relerr(getattr(se, k), getattr(so, k)) < 1e-10, "sink " + k
        o.evaluate(); e.evaluate()
        assert np.array_equal(o.tree()["order"], e.tree()["order"])
        do, de = o.diag(), e.diag()
        for k in do:
            assert relerr(de[k], do[k]) < 1e-10, k
