"""Device-side IC generator (sph_ics_disc; SURVEY.md 8(f)#2, replaces the sketch in Disc_ICs.py:1-41): the disc it
makes has the stated distribution, is reproducible from the seed, and any rank can make any slice of the rows."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H
from summersph_b200.ics import G_EFF
from test_multi_gpu import run_ranks

pytestmark = pytest.mark.gpu


def test_disc_statistics_and_determinism(built_engine):
    from summersph_b200.engine import Engine
    p = default_params(MODE_VARIABLE_H)
    n = 400_000
    with Engine(p) as e:
        e.ics_disc(n, seed=7)
        b, s = e.download()
        e.ics_disc(n, seed=7)
        b2, _ = e.download()
        e.ics_disc(n, seed=8)
        b3, _ = e.download()
    for k in ("x", "y", "z", "vx", "vy", "h"):
        assert np.array_equal(getattr(b, k), getattr(b2, k))
    assert not np.array_equal(b.x, b3.x)
    r = np.hypot(b.x, b.y)
    assert r.min() >= 10.0 and r.max() <= 100.0
    assert abs(np.mean(r * r) / ((100.0 ** 2 + 10.0 ** 2) / 2) - 1) < 5e-3          # uniform surface density: r^2 uniform
    v2 = b.vx ** 2 + b.vy ** 2
    assert np.max(np.abs(v2 * r / (G_EFF * 1.0) - 1)) < 1e-12 and np.all(b.vz == 0)   # Keplerian around the 1 M_sun sink
    assert np.max(np.abs(b.x * b.vx + b.y * b.vy)) < 1e-9                              # circular
    zeta = b.z / (0.05 * r)
    assert np.max(np.abs(zeta)) <= 3.0 + 1e-9 and abs(np.std(zeta) - 0.9973) < 0.005  # unit normal clamped at 3 sigma (as ics.keplerian_disc): variance 0.9707 + 18 (1 - Phi(3))
    assert abs(np.mean(zeta)) < 0.01
    assert abs(b.m.sum() - 0.01) < 1e-15 and np.all(b.u == 0.25) and np.all(b.alpha == 0.1)
    sigma = 0.01 / (np.pi * (100.0 ** 2 - 10.0 ** 2))
    rho = sigma / (np.sqrt(2 * np.pi) * 0.05 * r) * np.exp(-0.5 * zeta ** 2)
    assert np.max(np.abs(b.h / (1.2 * (b.m / rho) ** (1 / 3)) - 1)) < 1e-12
    assert len(s) == 1 and s.m[0] == 1.0 and s.radius[0] == p.sink_radius and s.x[0] == 0.0


def test_rank_slices_make_the_same_disc(built_engine, tmp_path):
    """Two domain-decomposed ranks, each generating only its own rows, hold the disc one rank generates."""
    one = run_ranks(tmp_path, "ics1", 1, "host", MODE_VARIABLE_H, steps=0, ics="devics")[0]
    two = run_ranks(tmp_path, "ics2", 2, "host", MODE_VARIABLE_H, steps=0, domains=1, ics="devics")
    for r in range(2):
        for k in ("x", "y", "z", "vx", "vy", "vz", "u", "m", "alpha", "h", "s_m"):
            assert np.array_equal(two[r][k], one[k]), (r, k)
