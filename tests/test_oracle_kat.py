"""Known-answer tests that pin the CPU oracle. The reference ships no tests or golden
vectors; these values follow directly from its formulas (SUMMER_SPH.f90:55-146, Variable.f90:119-141)."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, Bodies, Sinks, EVAL_ALL, EVAL_TREE, EVAL_DENSITY, EVAL_SPH, EVAL_GRAVITY, EVAL_SINKS
from oracle.oracle import Oracle


@pytest.fixture(scope="module")
def oF():
    return Oracle(default_params(MODE_FIXED_H))


@pytest.fixture(scope="module")
def oV():
    return Oracle(default_params(MODE_VARIABLE_H))


def test_effective_G(oF):
    assert oF.G == 39.478416442871094                      # real(4) literal, F:7
    assert float.hex(oF.G) == "0x1.3bd3cc0000000p+5"


def test_tables_at_sample_points(oF):
    w, dw, g = oF.tables()
    nq = 5000
    for q, ew, edw, eg in ((0.5, 0.71875, -0.9375, 0.136979166666667), (1.0, 0.25, -0.75, 19.0 / 30.0),
                           (1.5, 0.03125, -0.1875, 0.959895833333333), (2.0, 0.0, 0.0, 1.0)):
        i = int(round(q * nq / 2))
        assert w[i] == pytest.approx(ew, abs=1e-15)
        assert dw[i] == pytest.approx(edw, abs=1e-15)
        assert g[i] == pytest.approx(eg, abs=2e-15)
    assert w[0] == 1.0 and dw[0] == 0.0 and g[0] == 0.0


def test_lookup_kernel_fixed_h(oF):
    exp = {0.0: (0.02037183271576126, 0.0), 1.25: (0.014642254764453405, -0.0076394372684104725),
           2.5: (0.005092958178940315, -0.0061115498147283785), 3.75: (0.0006366197723675393, -0.0015278874536820946)}
    for r, (W, dW) in exp.items():
        gW, gdW = oF.lookup_kernel(r, 2.5)
        assert gW == pytest.approx(W, rel=1e-14, abs=1e-30)
        assert gdW == pytest.approx(dW, rel=1e-14, abs=1e-30)
    # q = 2 maps to i = nq-1, alpha ~ 1: tiny but not exactly zero (F:114)
    W, dW = oF.lookup_kernel(5.0, 2.5)
    assert 0.0 < W < 1e-24 and -1e-20 < dW < 0.0
    assert oF.lookup_kernel(5.0001, 2.5) == (0.0, 0.0)


def test_lookup_kernel_variable_h(oV):
    W, dW = oV.lookup_kernel(1.0, 0.7)                    # pi = real(4), V:7
    assert W == pytest.approx(0.04328948099846271, rel=1e-14)
    assert dW == pytest.approx(-0.32467084767397497, rel=1e-14)


def test_grav_kernel(oF):
    assert oF.lookup_grav_kernel(10.0, 2.5) == 1.0
    assert oF.lookup_grav_kernel(2.5, 2.5) == pytest.approx(19.0 / 30.0, rel=1e-13)


def _two(p, x2=(1.0, 0.5, 0.25), v2=(0.0, 0.0, 0.0), m=(1e-3, 2e-3), u=(1.0, 1.0), h=(1.0, 1.0), alpha=(1.0, 1.0)):
    b = Bodies([0.0, x2[0]], [0.0, x2[1]], [0.0, x2[2]], [0.0, v2[0]], [0.0, v2[1]], [0.0, v2[2]], list(u), list(m), list(alpha), list(h))
    return b, Sinks.empty(0)


def test_self_density(oF):
    # two far-apart particles: each sees only itself -> rho = m W(0) = m / (pi h^3)   (F:454, F:125)
    p = default_params(MODE_FIXED_H)
    b, s = _two(p, x2=(100.0, 0.0, 0.0))
    oF.upload(b, s); oF.evaluate(EVAL_TREE | EVAL_DENSITY)
    d = oF.diag()
    assert d["rho"][0] == pytest.approx(1e-3 / (3.14159265359 * 2.5 ** 3), rel=1e-14)
    assert d["rho"][1] == pytest.approx(2e-3 / (3.14159265359 * 2.5 ** 3), rel=1e-14)
    assert d["P"][0] == pytest.approx(0.4 * 1.0 * d["rho"][0], rel=1e-15)
    assert d["c"][0] == pytest.approx(np.sqrt(1.4 * 0.4), rel=1e-15)


def test_octant_rule_and_order(oF):
    # 8 particles, one per octant of a cube centred on the origin: child = 1 + [x>cx] + 2[y>cy] + 4[z>cz]
    # (strict >, F:208-214) and DFS order = children 1..8 (z most significant).
    pts = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], float)
    perm = np.array([5, 2, 7, 0, 3, 6, 1, 4])
    q = pts[perm]
    b = Bodies(q[:, 0], q[:, 1], q[:, 2], *[np.zeros(8)] * 3, np.ones(8), np.ones(8), np.zeros(8), np.ones(8))
    oF.upload(b, Sinks.empty(0)); oF.evaluate(EVAL_TREE)
    t = oF.tree()
    assert np.array_equal(perm[t["order"]], np.arange(8))
    assert np.all(t["level"] == 1) and np.all(t["size"] == 1.0)
    assert np.allclose(np.abs(t["cx"]), 0.5)


def test_pair_force_by_hand(oF):
    # F:356-391 evaluated by hand for one approaching pair
    p = default_params(MODE_FIXED_H)
    x2, v2 = (1.0, 0.5, 0.25), (-0.3, 0.1, 0.2)
    b, s = _two(p, x2=x2, v2=v2, alpha=(0.0, 0.0))
    oF.upload(b, s); oF.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_SPH)
    d = oF.diag()
    h = 2.5
    nr = np.array(x2); dr = np.sqrt(np.sum(nr ** 2)); vij = np.array(v2)
    W, dW = oF.lookup_kernel(dr, h)
    rho, P, c = d["rho"], d["P"], d["c"]
    # alpha is forced to 0 by upload (F:681) -> no viscosity
    gradW = nr / dr * dW
    A = (P[1] / rho[1] ** 2 + P[0] / rho[0] ** 2) * gradW
    assert np.allclose([d["ax"][1], d["ay"][1], d["az"][1]], -b.m[0] * A, rtol=1e-13)
    assert np.allclose([d["ax"][0], d["ay"][0], d["az"][0]], +b.m[1] * A, rtol=1e-13)
    vdg = float(gradW @ vij)
    assert d["udot"][1] == pytest.approx(b.m[0] * vdg * P[1] / rho[1] ** 2, rel=1e-13)
    assert d["udot"][0] == pytest.approx(b.m[1] * vdg * P[0] / rho[0] ** 2, rel=1e-13)
    lit = float(np.float32(0.15))
    adot1 = max(b.m[0] * vdg / rho[1], 0.0) + lit * ((0.1 - 0.0) * c[1] / h)
    assert d["alphadot"][1] == pytest.approx(adot1, rel=1e-13)


def test_pair_momentum_conservation(oV):
    from summersph_b200 import ics
    b, s = ics.keplerian_disc(800, seed=3)
    oV.upload(b, s); oV.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_SPH)
    d = oV.diag()
    for k in ("ax", "ay", "az"):
        tot = np.sum(b.m * d[k]); scale = np.sum(np.abs(b.m * d[k]))
        assert abs(tot) <= 1e-13 * scale                  # symmetric update F:383-384


def test_sink_gravity_unsoftened(oF):
    b = Bodies([10.0, -20.0], [0.0, 0.0], [0.0, 5.0], [0.0] * 2, [0.0] * 2, [0.0] * 2, [1.0] * 2, [1e-6, 2e-6], [0.0] * 2, [2.5] * 2)
    s = Sinks([0.0], [0.0], [0.0], [0.0], [0.0], [0.0], [1.0], [3.5])
    oF.upload(b, s); oF.evaluate(EVAL_TREE | EVAL_SINKS)
    d = oF.diag()
    G = oF.G
    r1 = np.array([10.0, 0.0, 0.0]); r2 = np.array([-20.0, 0.0, 5.0])
    a1 = -G * 1.0 * r1 / np.linalg.norm(r1) ** 3; a2 = -G * 1.0 * r2 / np.linalg.norm(r2) ** 3
    assert np.allclose([d["ax"][0], d["ay"][0], d["az"][0]], a1, rtol=1e-14)
    assert np.allclose([d["ax"][1], d["ay"][1], d["az"][1]], a2, rtol=1e-14)
    asink = G * (1e-6 * r1 / np.linalg.norm(r1) ** 3 + 2e-6 * r2 / np.linalg.norm(r2) ** 3)
    assert np.allclose([d["sink_ax"][0], d["sink_ay"][0], d["sink_az"][0]], asink, rtol=1e-13)


def test_dt_ladder(oF):
    # F:855-859 with real(4) bounds 0.1 and 0.0001
    b = Bodies([0.0, 1.0], [0.0] * 2, [0.0] * 2, [1.0, 1.0], [0.0] * 2, [0.0] * 2, [1.0] * 2, [1.0] * 2, [0.0] * 2, [2.5] * 2)
    oF.upload(b, Sinks.empty(0))
    # no evaluation: a = 0, udot = 0 -> t1 = inf, t2 = inf, t3 = h/|v| = 2.5, t4 = h/(2.2 c) with c = 0 -> inf
    assert oF.next_timestep(0.01) == 0.015                # cand = 0.625 > 2 dt and 1.5 dt < 0.1
    assert oF.next_timestep(0.06) == pytest.approx(0.09)
    assert oF.next_timestep(0.09) == 0.09                 # 1.5*0.09 = 0.135 > 0.1: stays
    assert oF.next_timestep(2.0) == 1.0                   # cand < 0.5 dt -> halve


def test_kepler_orbit_single_particle():
    # one light gas particle (+ a far, massless-ish companion so the tree has 2 leaves) round a 1 Msun sink
    p = default_params(MODE_FIXED_H)
    o = Oracle(p)
    G = o.G
    r0 = 30.0; v0 = np.sqrt(G / r0)
    b = Bodies([r0, 1400.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [v0, 0.0], [0.0, 0.0], [1e-8, 1e-8], [1e-20, 1e-20], [0.0] * 2, [2.5] * 2)
    s = Sinks([0.0], [0.0], [0.0], [0.0], [0.0], [0.0], [1.0], [3.5])
    o.upload(b, s)
    dt, t = 0.01, 0.0
    for _ in range(40):
        dt, t = o.step(dt, t)
    bb, _ = o.download()
    r = np.hypot(bb.x[0], bb.y[0])
    assert abs(r - r0) / r0 < 1e-4                        # KDK keeps the circular orbit
    E0 = 0.5 * v0 ** 2 - G / r0
    E = 0.5 * (bb.vx[0] ** 2 + bb.vy[0] ** 2) - G / r
    assert abs(E - E0) / abs(E0) < 1e-4
