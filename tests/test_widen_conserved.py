"""Conserved sums (`sph_conserved`, include/sph_b200.h): energy, momentum, angular momentum of the resident state.

The reference keeps no such bookkeeping; the definitions are this build's own.  CPU tests pin
the oracle's statement against direct numpy sums; the GPU tests compare the CUDA kernels with the oracle through
the C-ABI and check that asking for the sums never changes a later step.  Tolerances: sums of N same-sign terms
accumulated in a different order agree to ~sqrt(N) ulp, so 1e-10 relative to the sum of absolute terms."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SOFT_USES_HI, ics, Bodies, Sinks
from summersph_b200._abi import drift_report
from oracle.oracle import Oracle

TOL = 1e-10


def phi_ref(q):
    """Cubic-spline softened potential (per unit G M / h): the antiderivative of g(q)/q^2 with g from
    SUMMER_SPH.f90:91,94, continuous, -1/q beyond q = 2."""
    q = np.asarray(q, dtype=float)
    a = -1.4 + (2.0 / 3.0) * q ** 2 - 0.3 * q ** 4 + 0.1 * q ** 5
    with np.errstate(divide="ignore"):
        b = -1.6 + 1.0 / (15.0 * q) + (4.0 / 3.0) * q ** 2 - q ** 3 + 0.3 * q ** 4 - q ** 5 / 30.0
        c = -1.0 / q
    return np.where(q < 1.0, a, np.where(q < 2.0, b, c))


def direct_sums(b: Bodies, s: Sinks, G, h_of_i, soft_of_i):
    """O(N^2) numpy statement: every gas pair, i's own h and softening add-on (what a walk that opens every
    node down to the leaves sums)."""
    x = np.stack([b.x, b.y, b.z], 1); v = np.stack([b.vx, b.vy, b.vz], 1); m = b.m
    out = {"e_kin": 0.5 * np.sum(m * np.sum(v * v, 1)), "e_int": np.sum(m * b.u), "mass": np.sum(m)}
    P = np.sum(m[:, None] * v, 0); L = np.sum(m[:, None] * np.cross(x, v), 0)
    d2 = np.sum((x[:, None, :] - x[None, :, :]) ** 2, 2) + soft_of_i[:, None]
    dist = np.sqrt(d2)
    term = m[None, :] * phi_ref(dist / h_of_i[:, None]) / h_of_i[:, None]
    np.fill_diagonal(term, 0.0)
    out["e_pot_gas"] = 0.5 * G * np.sum(m * np.sum(term, 1))
    es = 0.0
    sx = np.stack([s.x, s.y, s.z], 1); sv = np.stack([s.vx, s.vy, s.vz], 1)
    for a in range(len(s)):
        if not s.m[a] > 0:
            continue
        es -= G * s.m[a] * np.sum(m / np.sqrt(np.sum((x - sx[a]) ** 2, 1)))
        out["e_kin"] += 0.5 * s.m[a] * np.sum(sv[a] ** 2); out["mass"] += s.m[a]
        P = P + s.m[a] * sv[a]; L = L + s.m[a] * np.cross(sx[a], sv[a])
        for c in range(a):
            if s.m[c] > 0:
                es -= G * s.m[a] * s.m[c] / np.sqrt(np.sum((sx[a] - sx[c]) ** 2))
    out["e_pot_sink"] = es
    out["e_pot"] = out["e_pot_gas"] + es
    out.update(px=P[0], py=P[1], pz=P[2], lx=L[0], ly=L[1], lz=L[2])
    return out


def two_sinks():
    s = Sinks.empty(2)
    s.x[:] = [0.0, 30.0]; s.y[:] = [0.0, 5.0]; s.vy[:] = [0.0, 1.5]; s.m[:] = [1.0, 0.02]; s.radius[:] = [0.5, 0.5]
    return s


def test_soft_potential_is_the_antiderivative_of_the_force_table():
    """d phi / dq = g(q)/q^2 against the reference's grav_table (F:81-101) at the table nodes."""
    o = Oracle(default_params(MODE_FIXED_H))
    _, _, g = o.tables()
    nq = 5000; q = np.arange(nq + 1) * (2.0 / nq)
    dq = 1e-6
    num = (phi_ref(q[5:] + dq) - phi_ref(q[5:] - dq)) / (2 * dq)
    assert np.max(np.abs(num[:-1] - g[5:-1] / q[5:-1] ** 2)) < 1e-8
    assert phi_ref(0.0) == -1.4 and abs(phi_ref(2.0) + 0.5) < 1e-15 and abs(phi_ref(1.0 - 1e-14) - phi_ref(1.0)) < 1e-12


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI])
def test_oracle_conserved_matches_direct_sums(mode):
    """theta -> 0 opens every node down to the single-particle leaves: the tree sum is the pair sum."""
    p = default_params(mode, theta=1e-12, theta_override=1)
    b, _ = ics.keplerian_disc(700, seed=5)
    s = two_sinks()
    o = Oracle(p); o.upload(b, s)
    got = o.conserved()
    h = b.h if mode & MODE_VARIABLE_H else np.full(len(b), p.h_fixed)
    soft = 0.001 * b.h if mode & FLAG_SOFT_USES_HI else np.full(len(b), 0.001 * p.h_fixed)
    ref = direct_sums(b, s, o.G, h, soft)
    scale_p = np.sqrt(2 * ref["e_kin"] * ref["mass"])
    for k in ("e_kin", "e_int", "e_pot_gas", "e_pot_sink", "e_pot", "mass", "lz"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]), k
    for k in ("px", "py", "pz"):
        assert abs(got[k] - ref[k]) <= 1e-12 * scale_p, k
    for k in ("lx", "ly"):
        assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref["lz"]), k
    assert got["e_total"] == got["e_kin"] + got["e_int"] + got["e_pot"]


def test_oracle_conserved_bh_close_to_direct():
    """theta = 0.5 monopoles: the tree potential stays within the Barnes-Hut error of the pair sum (Appendix D)."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(1500, seed=9)
    o = Oracle(p); o.upload(b, s)
    got = o.conserved()
    ref = direct_sums(b, Sinks.empty(0), o.G, b.h, np.full(len(b), 0.001 * p.h_fixed))
    assert abs(got["e_pot_gas"] - ref["e_pot_gas"]) < 2e-2 * abs(ref["e_pot_gas"])
    # the central sink dominates: E_pot_sink = -G M_star sum m/r
    r = np.sqrt(b.x ** 2 + b.y ** 2 + b.z ** 2)
    assert got["e_pot_sink"] == pytest.approx(-o.G * 1.0 * np.sum(b.m / r), rel=1e-12)


def test_no_sink_file_has_no_sink_terms():
    """The reference's dummy zero-mass sink (F:698-707) contributes nothing, even with a particle at the origin."""
    b, _ = ics.uniform_sphere(300)
    b.x[0] = b.y[0] = b.z[0] = 0.0
    o = Oracle(default_params(MODE_VARIABLE_H)); o.upload(b, Sinks.empty(0))
    got = o.conserved()
    assert got["e_pot_sink"] == 0.0 and np.isfinite(got["e_total"]) and got["e_kin"] == 0.0


def test_drift_report_kepler_orbit_short_run():
    """A cold, light disc around the sink for a few steps: the leapfrog keeps E and L to the dt^2 level."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(400, seed=3, m_disc=1e-6)
    o = Oracle(p); o.upload(b, s)
    first = o.conserved()
    dt, t = 0.01, 0.0
    for _ in range(3):
        dt, t = o.step(dt, t)
    rep = drift_report(first, o.conserved())
    assert abs(rep["energy_rel"]) < 1e-3 and rep["angular_momentum_rel"] < 1e-3 and rep["mass_rel"] == 0.0


def test_simulate_reports_drift_only_when_asked():
    """The host shell (simulate) logs the reference's per-step line only; a `drift` dict adds the report."""
    from summersph_b200.simulate import simulate
    p = default_params(MODE_VARIABLE_H, end_time=0.03)
    b, s = ics.keplerian_disc(300, seed=8, m_disc=1e-6)
    s.radius[:] = p.sink_radius
    plain, logged, rep = [], [], {}
    simulate(b, s, p, engine=Oracle(p), log=plain.append)
    simulate(b, s, p, engine=Oracle(p), log=logged.append, drift=rep)
    assert all(l.startswith(" SPH Particles:") for l in plain)
    assert logged[:len(plain)] == plain and len(logged) == len(plain) + 2 and logged[-1].startswith(" Drift over")
    assert rep["steps"] == len(plain) and abs(rep["energy_rel"]) < 1e-3 and rep["first"]["mass"] == rep["last"]["mass"]


# ------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels against the oracle, through the C-ABI
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def E(built_engine):
    from summersph_b200.engine import Engine
    return Engine


def assert_same_sums(ge, go):
    abs_scale = {k: max(abs(go[k]), 1e-300) for k in ("e_kin", "e_int", "e_pot", "e_pot_gas", "e_pot_sink", "mass")}
    for k, sc in abs_scale.items():
        if k == "e_pot_sink" and go[k] == 0.0:
            assert ge[k] == 0.0
            continue
        assert abs(ge[k] - go[k]) <= TOL * sc, (k, ge[k], go[k])
    p_scale = np.sqrt(2 * abs(go["e_kin"]) * go["mass"]) or 1.0
    l_scale = max(np.sqrt(go["lx"] ** 2 + go["ly"] ** 2 + go["lz"] ** 2), 1e-300)
    for k in ("px", "py", "pz"):
        assert abs(ge[k] - go[k]) <= TOL * p_scale, (k, ge[k], go[k])
    for k in ("lx", "ly", "lz"):
        assert abs(ge[k] - go[k]) <= TOL * l_scale, (k, ge[k], go[k])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI])
def test_gpu_conserved_matches_oracle(mode, E):
    p = default_params(mode)
    b, s = ics.keplerian_disc(10_000)
    o = Oracle(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        assert_same_sums(e.conserved(), o.conserved())        # straight after the upload: builds its own tree
        dto = dte = 0.01; to = te = 0.0
        for _ in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
        assert (dto, to) == (dte, te)
        assert_same_sums(e.conserved(), o.conserved())        # after a step: the tree of evaluation B is reused


@pytest.mark.gpu
def test_gpu_conserved_two_sinks_and_removals(E):
    """Accreting sinks and a tight bounding cube (the set-up of test_gpu_parity.test_accretion_and_bounds):
    particles leave, so the sums need a fresh tree after every step."""
    p = default_params(MODE_VARIABLE_H, bounding_size=85.0)
    b, _ = ics.keplerian_disc(8_000, seed=9)
    s = Sinks([0.0, 40.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 6.0], [0.0, 0.0], [1.0, 0.01], [13.0, 6.0])
    o = Oracle(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert o.sizes() == e.sizes()
            assert_same_sums(e.conserved(), o.conserved())
        assert o.sizes()[0] < 8_000


@pytest.mark.gpu
def test_gpu_conserved_depth_limited_tree(E):
    """max_depth below the natural leaf depth: gravity takes the multi-particle childless nodes whole (F:182),
    and so does the potential (a particle's own depth-limited node is not a single-particle leaf: kept)."""
    p = default_params(MODE_VARIABLE_H, max_depth=5)
    b, s = ics.keplerian_disc(6_000, seed=4)
    o = Oracle(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        assert_same_sums(e.conserved(), o.conserved())


@pytest.mark.gpu
def test_gpu_conserved_does_not_change_the_run(E):
    """Asking for the sums between steps (which may build the tree early) leaves every later step bit-identical."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(8000, seed=21)
    out = []
    for ask in (False, True):
        with E(p) as e:
            e.upload(b, s)
            dt, t = 0.01, 0.0
            if ask:
                e.conserved()
            for _ in range(3):
                dt, t = e.step(dt, t)
                if ask:
                    e.conserved()
            out.append((dt, t) + e.download())
    (d1, t1, b1, s1), (d2, t2, b2, s2) = out
    assert (d1, t1) == (d2, t2)
    for k in ("x", "y", "z", "vx", "vy", "vz", "u", "alpha", "h"):
        assert np.array_equal(getattr(b1, k), getattr(b2, k)), k
    assert np.array_equal(s1.x, s2.x) and np.array_equal(s1.vx, s2.vx)
