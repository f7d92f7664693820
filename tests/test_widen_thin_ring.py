"""BASELINE.json configs[2] / SURVEY.md §8(d) cfg 3: thin ring around a 1 M_sun sink, variable h, tree gravity —
"run >= 20 steps; compare Sigma(r, t) spreading".

GPU: (a) 20 loop bodies of a 10k ring on the CUDA engine and on the oracle: dt / t equal step by step, the states agree
to 1e-10 and the surface-density profiles Sigma(r) binned from them are the same; (b) the full 1M-particle ring for 20
steps on the engine alone (the oracle would need hours): nothing is lost, energy and angular momentum drift stay small,
the ring stays where it was and does not contract (viscous spreading is outward / inward symmetric at this order).
CPU: the Sigma(r) estimator itself."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, ics
from summersph_b200._abi import drift_report
from summersph_b200.state import GAS_FIELDS
from conftest import relerr

N_SMALL, N_FULL, STEPS = 10_000, 1_000_000, 20
R_EDGES = np.linspace(35.0, 65.0, 61)


def surface_density(b, edges=R_EDGES):
    """Sigma(r) = mass in the annulus / its area, from the particles' cylindrical radii."""
    r = np.hypot(b.x, b.y)
    msum, _ = np.histogram(r, bins=edges, weights=b.m)
    return msum / (np.pi * (edges[1:] ** 2 - edges[:-1] ** 2))


def ring_moments(b):
    r = np.hypot(b.x, b.y)
    mean = np.sum(b.m * r) / np.sum(b.m)
    return mean, np.sqrt(np.sum(b.m * (r - mean) ** 2) / np.sum(b.m))


def test_surface_density_estimator():
    b, s = ics.thin_ring(200_000, seed=3)
    sig = surface_density(b)
    area = np.pi * (R_EDGES[1:] ** 2 - R_EDGES[:-1] ** 2)
    assert np.sum(sig * area) == pytest.approx(np.sum(b.m), rel=1e-12)          # every particle lies in 35 < r < 65
    mean, width = ring_moments(b)
    assert mean == pytest.approx(50.0, abs=0.05) and width == pytest.approx(2.5, rel=0.02)
    rc = 0.5 * (R_EDGES[1:] + R_EDGES[:-1])
    model = 1e-3 / (2 * np.pi * rc) / (np.sqrt(2 * np.pi) * 2.5) * np.exp(-0.5 * ((rc - 50.0) / 2.5) ** 2)
    core = np.abs(rc - 50.0) < 5.0
    assert np.max(np.abs(sig[core] / model[core] - 1.0)) < 0.1


@pytest.fixture(scope="module")
def E(built_engine):
    from summersph_b200.engine import Engine
    return Engine


@pytest.mark.gpu
def test_ring_20_steps_engine_follows_oracle(E):
    from oracle.oracle import Oracle
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.thin_ring(N_SMALL)
    o = Oracle(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for k in range(STEPS):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert (dto, to) == (dte, te), k
        be, se = e.download(); bo, so = o.download()
    assert len(be) == len(bo) == N_SMALL
    for k in GAS_FIELDS:
        assert relerr(getattr(be, k), getattr(bo, k)) < 1e-10, k
    sig_e, sig_o = surface_density(be), surface_density(bo)
    assert np.max(np.abs(sig_e - sig_o)) <= 1e-9 * np.max(sig_o)
    assert to > 0.5                                                              # the dt ladder opened up: a real stretch of time


@pytest.mark.gpu
def test_ring_1M_20_steps(E):
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.thin_ring(N_FULL)
    mean0, width0 = ring_moments(b)
    sig0 = surface_density(b)
    with E(p) as e:
        e.upload(b, s)
        first = e.conserved()
        dt, t = 0.01, 0.0
        for _ in range(STEPS):
            dt, t = e.step(dt, t)
        last = e.conserved()
        b1, s1 = e.download()
    rep = drift_report(first, last)
    print({"t": t, "dt": dt, **rep})
    assert len(b1) == N_FULL and len(s1) == 1 and rep["mass_rel"] == 0.0
    assert abs(rep["energy_rel"]) < 1e-3 and rep["angular_momentum_rel"] < 1e-6
    mean1, width1 = ring_moments(b1)
    assert mean1 == pytest.approx(mean0, abs=0.05) and width1 > 0.98 * width0 and width1 < 1.2 * width0
    sig1 = surface_density(b1)
    area = np.pi * (R_EDGES[1:] ** 2 - R_EDGES[:-1] ** 2)
    assert np.sum(sig1 * area) == pytest.approx(np.sum(sig0 * area), rel=1e-6)   # the ring is still inside 35 < r < 65
    assert np.all(np.isfinite(b1.u)) and np.all(b1.h > 0)
