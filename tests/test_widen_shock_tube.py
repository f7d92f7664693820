"""BASELINE.json configs[1]: Sod shock tube, 100k particles, artificial viscosity on, "compared against the analytic
solution and the reference".

CPU: the exact Riemann solver (summersph_b200/analytic.py) against Toro's published star values for Sod's problem and
against the conservation laws; the tube generator's geometry; the Omega ~ 3 of the variable-h program on a lattice
and the effective gas that follows from it.  GPU: the CUDA engine runs both programs' 100k tubes with the reference's
own loop and its L1 errors against the exact solutions must stay within bounds taken from measured runs
(profiles/r1_sod_100k.md lists engine and oracle side by side).
"""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, ics
from summersph_b200.analytic import (riemann_star, riemann_exact, sod_exact, sod_core_mask, effective_gas, SOD,
                                     OMEGA_LATTICE)


def test_sod_star_region_matches_published_values():
    """Toro, Table 4.2, Test 1 (gamma = 1.4): p* = 0.30313, u* = 0.92745, rho*L = 0.42632, rho*R = 0.26557."""
    ps, us = riemann_star(**SOD)
    assert ps == pytest.approx(0.30313, abs=5e-6) and us == pytest.approx(0.92745, abs=5e-6)
    rho, v, p = sod_exact(np.array([-0.4, 0.1, 0.3, 0.45]), 0.25)
    assert rho[0] == 1.0 and v[0] == 0.0 and p[0] == 1.0
    assert rho[1] == pytest.approx(0.42632, abs=5e-6) and rho[2] == pytest.approx(0.26557, abs=5e-6)
    assert rho[3] == 0.125 and p[3] == 0.1 and v[3] == 0.0
    # shock speed from the Rankine-Hugoniot mass flux: s = rho* u* / (rho* - rho_R)
    s = rho[2] * us / (rho[2] - 0.125)
    assert s == pytest.approx(1.75216, abs=5e-5)
    r_edge, _, _ = sod_exact(np.array([s * 0.25 - 1e-9, s * 0.25 + 1e-9]), 0.25)
    assert r_edge[0] == pytest.approx(0.26557, abs=5e-6) and r_edge[1] == 0.125


def test_exact_solution_obeys_the_conservation_laws():
    """Integrals of the sampled solution over a box the waves have not left: mass constant, momentum grows as
    t (p_L - p_R), total energy constant."""
    t, g = 0.2, 1.4
    x = np.linspace(-0.5, 0.5, 400_001)
    rho, v, p = sod_exact(x, t)
    dx = x[1] - x[0]
    integ = lambda f: float(np.sum(0.5 * (f[1:] + f[:-1])) * dx)   # noqa: E731
    assert integ(rho) == pytest.approx(0.5 * 1.0 + 0.5 * 0.125, abs=2e-6)
    assert integ(rho * v) == pytest.approx(t * (1.0 - 0.1), abs=2e-6)
    assert integ(0.5 * rho * v * v + p / (g - 1)) == pytest.approx(0.5 * 1.0 / 0.4 + 0.5 * 0.1 / 0.4, abs=5e-6)


def test_riemann_solver_handles_two_shocks_and_two_rarefactions():
    """Symmetric collisions / expansions: u* = 0 by symmetry; both branches of each side are exercised."""
    ps, us = riemann_star(1.0, 2.0, 1.0, 1.0, -2.0, 1.0)
    assert us == pytest.approx(0.0, abs=1e-12) and ps > 1.0
    rho, v, p = riemann_exact(np.array([-3.0, 0.0, 3.0]), 1.0, 2.0, 1.0, 1.0, -2.0, 1.0)
    assert rho[1] > 1.0 and v[1] == pytest.approx(0.0, abs=1e-12) and rho[0] == rho[2] == 1.0
    ps, us = riemann_star(1.0, -0.5, 1.0, 1.0, 0.5, 1.0)
    assert us == pytest.approx(0.0, abs=1e-12) and ps < 1.0
    rho, v, p = riemann_exact(np.array([-3.0, -0.1, 0.1, 3.0]), 1.0, -0.5, 1.0, 1.0, 0.5, 1.0)
    assert rho[1] == pytest.approx(rho[2]) and rho[1] < 1.0 and v[0] == -0.5 and v[3] == 0.5
    assert p[1] / rho[1] ** 1.4 == pytest.approx(1.0, rel=1e-12)     # rarefactions are isentropic


def test_sod_box_geometry():
    b, s, g = ics.sod_box(20_000, 0.2, rho_scale=1e-9)
    assert len(s) == 0 and abs(len(b) - 20_000) < 2_000
    assert np.all(b.m == b.m[0]) and np.all(b.vx == 0.0)
    left = b.x < 0
    assert b.x.min() == pytest.approx(-g["len_l"] + 0.5 * g["dl"]) and b.x.max() == pytest.approx(g["len_r"] - 0.5 * g["dr"])
    # lattice densities 1 : 0.125 and pressures 1 : 0.1 (P = (gamma-1) rho u)
    assert b.m[0] / g["dl"] ** 3 == pytest.approx(1e-9) and b.m[0] / g["dr"] ** 3 == pytest.approx(0.125e-9)
    assert 0.4 * 1.0 * b.u[left][0] == pytest.approx(1.0) and 0.4 * 0.125 * b.u[~left][0] == pytest.approx(0.1)
    assert np.min(b.x ** 2 + b.y ** 2 + b.z ** 2) > 0.0                # the dummy sink sits at the origin (F:698-707)
    assert np.all(b.h[left] == pytest.approx(1.2 * g["dl"])) and b.h.min() > 0.01   # V:528: h <= 0.01 never updates
    # at t_end the core still holds the whole wave pattern and at least core_cells^2 columns of right-hand cells
    m = sod_core_mask(b.x, b.y, b.z, g["t_end"], g)
    assert b.x[m].min() < -1.1832 * 0.2 and b.x[m].max() > 1.7522 * 0.2
    cols = np.unique(np.round(np.stack([b.y[m & ~left], b.z[m & ~left]], 1) / g["dr"], 3), axis=0)
    assert len(cols) >= 16
    with pytest.raises(ValueError):
        sod_core_mask(b.x, b.y, b.z, 5 * g["t_end"], g)


def test_variable_h_program_has_omega_3_on_a_lattice():
    """V:455,487 as coded give Omega = 1 + (1/(3 rho)) sum m (3W - r dW/dr) ~ 1 + 2 on a uniform lattice (the textbook
    grad-h factor is ~ 1): every pressure term of that program is divided by ~3.  The constant the tube sizing uses."""
    from oracle.oracle import Oracle
    from summersph_b200 import EVAL_TREE, EVAL_DENSITY
    b, s, g = ics.sod_box(20_000, 0.2, rho_scale=1e-9)
    o = Oracle(default_params(MODE_VARIABLE_H), threads=4); o.upload(b, s); o.evaluate(EVAL_TREE | EVAL_DENSITY)
    om = o.diag()["omega"][sod_core_mask(b.x, b.y, b.z, 0.0, g)]
    assert np.median(om) == pytest.approx(OMEGA_LATTICE, abs=0.005) and om.std() < 0.2


def test_effective_gas_of_the_variable_h_program():
    g_eff, st = effective_gas(1.4, 3.0)
    assert g_eff == pytest.approx(1.0 + 0.4 / 3.0) and st["p_l"] == pytest.approx(1.0 / 3.0) and st["rho_r"] == 0.125
    # same specific energies as the ICs: u = P_eff / ((gamma_eff - 1) rho) = P / ((gamma - 1) rho)
    assert st["p_l"] / ((g_eff - 1.0) * st["rho_l"]) == pytest.approx(2.5) and st["p_r"] / ((g_eff - 1.0) * st["rho_r"]) == pytest.approx(2.0)
    assert effective_gas(1.4, 1.0) == (1.4, SOD)
    # slower waves: sound speed of the left state drops by sqrt(gamma_eff / (gamma Omega))
    ps, vs = riemann_star(gamma=g_eff, **st)
    assert vs < 0.7 and np.sqrt(g_eff * st["p_l"]) == pytest.approx(np.sqrt(1.4) * np.sqrt(g_eff / (1.4 * 3.0)))
    # a tube sized for them is smaller and finer at the same particle count
    _, _, g1 = ics.sod_box(50_000, 0.2); _, _, g3 = ics.sod_box(50_000, 0.2, omega=3.0)
    assert g3["dl"] < 0.7 * g1["dl"] and g3["width"] < g1["width"]


# ------------------------------------------------------------------------------------------------------
# Bounds for the engine at 100k particles (~11 right-hand spacings of shock travel): about 1.3 x the L1 errors the
# engine reached on a B200 in round 1 (profiles/r1_sod_100k.md; the oracle's 20k runs sit beside them there).
#   fixed h    vs Sod, gamma = 1.4        : rho 0.045  v 0.051  u 0.029
#   variable h vs the effective gas       : rho 0.095  v 0.126  u 0.019     (Omega0 = 2.98: gamma_eff 1.134, P/2.98)
#   variable h vs Sod, gamma = 1.4        : rho 0.258  v 0.435  u 0.235     (the reference's Omega: waves ~0.6 x slower)
BOUNDS = {"fixed": {"rho_l1": 0.06, "v_l1": 0.07, "u_l1": 0.04},
          "variable": {"rho_l1": 0.125, "v_l1": 0.165, "u_l1": 0.03}}


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["fixed", "variable"])
def test_sod_100k_against_the_exact_solution(kind, built_engine):
    from summersph_b200.engine import Engine
    from sod_report import run_sod, sod_case
    p, b, s, geom = sod_case(kind, 100_000)
    with Engine(p) as e:
        rep = run_sod(e, b, s, geom)
    print(rep)
    assert rep["n_core"] > 1000 and rep["t"] >= geom["t_end"]
    which = "nominal" if kind == "fixed" else "effective"
    for k, bound in BOUNDS[kind].items():
        assert rep[which][k] < bound, (k, rep[which][k])
    if kind == "variable":
        assert rep["omega0"] == pytest.approx(OMEGA_LATTICE, abs=0.005)
        # the deviation from the textbook gas is the reference's, and it is large: keep it visible
        assert rep["nominal"]["rho_l1"] > 2.0 * rep["effective"]["rho_l1"] and rep["nominal"]["v_l1"] > 0.3
    # no sinks, nothing leaves the bounding cube: mass exact, total energy to the integrator's accuracy,
    # momentum and angular momentum of the symmetric pair forces to rounding (tree gravity is ~1e-9 of the forces here)
    d = rep["drift"]
    assert d["mass_rel"] == 0.0 and abs(d["energy_rel"]) < 1e-2
    assert d["momentum_rel"] < 1e-6
