"""Parity of the CUDA engine (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bit-exact: Morton (DFS leaf) order, leaf levels and cells, neighbour sets, dt ladder, interaction counts.
FP64 fields: 1e-10 relative (north_star), scale = max(|value|, field RMS) — SURVEY.md §8(c)."""
import numpy as np
import pytest

from summersph_b200 import (default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SOFT_USES_HI, ics, Bodies, Sinks,
                            EVAL_ALL, EVAL_TREE, EVAL_DENSITY, EVAL_GRAVITY, EVAL_SINKS, EVAL_SPH)
from summersph_b200.state import GAS_FIELDS
from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def E(built_engine):
    from summersph_b200.engine import Engine
    return Engine


@pytest.fixture(scope="module")
def O():
    from oracle.oracle import Oracle
    return Oracle


def compare_eval(o, e, check_ngb=True, tol=TOL, mask=EVAL_ALL):
    """`e` was just evaluated with `mask` in its default mode (exact-zero pairs culled before the box test)."""
    to, te = o.tree(), e.tree()
    assert np.array_equal(to["order"], te["order"])
    assert np.array_equal(to["level"], te["level"])
    for k in ("cx", "cy", "cz", "size"):
        assert np.array_equal(to[k], te[k]), k
    if check_ngb:
        co, ho, _, _ = o.neighbours(with_list=False); ce, he, _, _ = e.neighbours(with_list=False)
        assert np.array_equal(co, ce) and np.array_equal(ho, he)
    do, de = o.diag(), e.diag()
    for k in do:
        assert relerr(de[k], do[k]) < tol, f"{k}: {relerr(de[k], do[k]):.3e}"
    # the reference's interaction counts need every leaf-box candidate: evaluate again with exact counters.
    # The culled pairs only ever add exact zeros; the tiles they no longer occupy change the order in which
    # the two interleaved density chains are summed, so the fields agree to rounding, not bit for bit.
    e.set_exact_counters(True); e.evaluate(mask); dx = e.diag(); e.set_exact_counters(False)
    for k in de:
        assert relerr(dx[k], de[k]) < 1e-13, k
        assert relerr(dx[k], do[k]) < tol, k
    co, ce = o.counters(), e.counters()
    for k in ("density_candidates", "density_contributing", "sph_pairs", "grav_accepted"):
        assert co[k] == ce[k], k


def compare_state(o, e, tol=TOL):
    bo, so = o.download(); be, se = e.download()
    assert len(bo) == len(be) and len(so) == len(se)
    for k in GAS_FIELDS:
        assert relerr(getattr(be, k), getattr(bo, k)) < tol, f"{k}: {relerr(getattr(be, k), getattr(bo, k)):.3e}"
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
        assert relerr(getattr(se, k), getattr(so, k)) < tol, f"sink {k}"


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI])
def test_disc_10k_evaluation(mode, E, O):
    """config 1: Keplerian disc, 10k gas + 1 central sink."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(10_000)
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
    with E(p) as e:
        e.upload(b, s); e.evaluate()
        compare_eval(o, e)
        # neighbour lists row by row
        _, _, oo, lo = o.neighbours(); _, _, oe, le = e.neighbours()
        assert np.array_equal(oo, oe) and np.array_equal(lo, le)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_disc_10k_phases(mode, E, O):
    p = default_params(mode)
    b, s = ics.keplerian_disc(10_000, seed=77)
    o = O(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        for mask in (EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY, EVAL_TREE | EVAL_DENSITY | EVAL_SINKS, EVAL_TREE | EVAL_DENSITY | EVAL_SPH):
            o.evaluate(mask); e.evaluate(mask)
            do, de = o.diag(), e.diag()
            for k in do:
                assert relerr(de[k], do[k]) < TOL, (mask, k)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_disc_10k_steps(mode, E, O):
    """config 1 to its end time (0.1 yr ~ 6 steps on the dt ladder): state, dt and t track the oracle."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(10_000)
    o = O(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        while to < 0.1:
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert dto == dte and to == te
            assert o.sizes() == e.sizes()
            if mode & MODE_VARIABLE_H:
                assert o.counters()["h_iterations"] == e.counters()["h_iterations"]
        compare_state(o, e)


def test_sod_tube_variable(E, O):
    """config 2 (reduced to ~20k for the CPU oracle): lattice ICs, no sink -> dummy sink, alpha = 1."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.sod_tube(20_000, width=6)
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
    with E(p) as e:
        e.upload(b, s); e.evaluate()
        compare_eval(o, e)
        o.record_neighbours(False)
        o.upload(b, s); e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(3):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert dto == dte and to == te
        compare_state(o, e)


def test_sod_tube_fixed(E, O):
    b, s = ics.sod_tube(20_000, width=6)
    p = default_params(MODE_FIXED_H, h_fixed=float(1.2 * np.max(b.h) / 1.2))
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
    with E(p) as e:
        e.upload(b, s); e.evaluate()
        compare_eval(o, e)


def test_thin_ring_50k(E, O):
    """config 3 geometry at 50k (the oracle's comfortable range), variable h, 2 steps."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.thin_ring(50_000)
    o = O(p, threads=1); o.record_neighbours(True); o.upload(b, s); o.evaluate()
    with E(p) as e:
        e.upload(b, s); e.evaluate()
        compare_eval(o, e)
        o.record_neighbours(False)
        o.upload(b, s); e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert dto == dte and to == te
        compare_state(o, e)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_accretion_and_bounds(mode, E, O):
    """Sinks that accrete (two of them, overlapping reach) and a tight bounding cube: removals, renumbering,
    sink mass/position/velocity updates and the order-preserving pack (F:484-556 | V:616-688, F:471-482)."""
    p = default_params(mode, bounding_size=85.0)
    b, _ = ics.keplerian_disc(8_000, seed=9)
    s = Sinks([0.0, 40.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 6.0], [0.0, 0.0], [1.0, 0.01], [13.0, 6.0])
    o = O(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for k in range(3):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert o.sizes() == e.sizes(), k
            assert dto == dte and to == te
        assert o.sizes()[0] < 8_000
        compare_state(o, e)


def test_sink_creation(E, O):
    """V:549-597: the first over-dense particle (m (eta/h)^3 > 0.5) far from every sink spawns a sink
    (mass 1e-11, radius 2h), which then accretes its own seed particle in the same step (V:1155-1157).
    max_length below the Newton-Raphson proposal keeps the seed's h (V:528,541)."""
    p = default_params(MODE_VARIABLE_H, max_length=0.21)
    b, s = ics.keplerian_disc(4_000, seed=21)
    b.m[1234] = 5e-3; b.h[1234] = 0.2
    o = O(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto, to = o.step(0.01, 0.0); dte, te = e.step(0.01, 0.0)
        assert o.sizes() == e.sizes() == (3999, 2)
        assert (dto, to) == (dte, te)
        compare_state(o, e)
        dto, to = o.step(dto, to); dte, te = e.step(dte, te)
        assert o.sizes() == e.sizes()
        compare_state(o, e)


@pytest.mark.parametrize("max_depth", [4, 7])
def test_depth_limited_tree(max_depth, E, O):
    """max_depth below the natural leaf depth: multi-particle childless nodes are skipped by the density /
    SPH walks and taken whole by gravity (F:182,431,443)."""
    p = default_params(MODE_VARIABLE_H, max_depth=max_depth)
    b, s = ics.keplerian_disc(6_000, seed=4)
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY | EVAL_SINKS)
    with E(p) as e:
        e.upload(b, s); e.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY | EVAL_SINKS)
        to, te = o.tree(), e.tree()
        assert np.array_equal(to["level"], te["level"]) and np.max(te["level"]) == max_depth
        assert np.any(to["n_in_leaf"] > 1)
        co, ho, _, _ = o.neighbours(with_list=False); ce, he, _, _ = e.neighbours(with_list=False)
        assert np.array_equal(co, ce) and np.array_equal(ho, he)
        do, de = o.diag(), e.diag()
        ok = do["rho"] > 0
        assert relerr(de["rho"][ok], do["rho"][ok]) < TOL
        assert np.array_equal(de["rho"] == 0, do["rho"] == 0)
        for k in ("ax", "ay", "az"):
            assert relerr(de[k], do[k]) < TOL, k
        # SPH pass: particles in a multi-particle node are found by nobody (rho = 0 -> P/rho^2 = NaN) but still
        # visit their lower-numbered neighbours and hand them that NaN (F:354-391): same NaN pattern, same finite values
        o.evaluate(); e.evaluate()
        do, de = o.diag(), e.diag()
        for k in ("ax", "udot", "alphadot"):
            nan_o, nan_e = np.isnan(do[k]), np.isnan(de[k])
            assert np.array_equal(nan_o, nan_e), k
            assert not nan_o.all()
            fin = ~nan_o
            assert relerr(de[k][fin], do[k][fin]) < TOL, k


def test_two_word_keys(E, O):
    """Pairs closer than root_size/2^21 overflow the 63-bit key: the engine switches to two-word (42-level)
    keys and still reproduces the reference tree (leaf levels > 21) and everything downstream."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(5_000, seed=31)
    for a_, c_, eps in ((10, 11, 3e-7), (200, 201, 5e-9), (3000, 3001, 2e-10)):
        b.x[c_] = b.x[a_] + eps; b.y[c_] = b.y[a_] - 0.5 * eps; b.z[c_] = b.z[a_] + 0.25 * eps
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
    assert np.max(o.tree()["level"]) > 21
    with E(p) as e:
        e.upload(b, s); e.evaluate()
        compare_eval(o, e)
        o.record_neighbours(False)
        o.upload(b, s); e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(2):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert dto == dte and to == te
        compare_state(o, e)


def test_depth_limit_between_one_and_two_words(E, O):
    """max_depth = 30 with coincident particles: a depth-limited multi-particle leaf at level 30."""
    p = default_params(MODE_VARIABLE_H, max_depth=30)
    b, s = ics.keplerian_disc(3_000, seed=32)
    for k in ("x", "y", "z"):
        getattr(b, k)[7] = getattr(b, k)[6]
    mask = EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY | EVAL_SINKS
    o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate(mask)
    with E(p) as e:
        e.upload(b, s); e.evaluate(mask)
        to, te = o.tree(), e.tree()
        assert np.array_equal(to["order"], te["order"]) and np.array_equal(to["level"], te["level"])
        assert te["level"][6] == 30 and te["level"][7] == 30
        co, ho, _, _ = o.neighbours(with_list=False); ce, he, _, _ = e.neighbours(with_list=False)
        assert np.array_equal(co, ce) and np.array_equal(ho, he)
        do, de = o.diag(), e.diag()
        for k in ("ax", "ay", "az"):
            assert relerr(de[k], do[k]) < TOL, k


def test_key_depth_error_is_loud(E):
    """Two coincident particles with max_depth = 1000: not even the 126-bit key separates them -> SPH_ERR_DEPTH."""
    from summersph_b200.engine import SphError
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(2_000, seed=2)
    for k in ("x", "y", "z"):
        getattr(b, k)[1] = getattr(b, k)[0]
    with E(p) as e:
        e.upload(b, s)
        with pytest.raises(SphError) as ei:
            e.evaluate()
        assert ei.value.code == -6


def test_ragged_and_tiny_inputs(E, O):
    """Sizes that do not fill a warp / a walk group, and the smallest legal input (2 particles)."""
    from summersph_b200.engine import SphError
    p = default_params(MODE_VARIABLE_H)
    for n in (2, 3, 31, 33, 65, 1000):
        b, s = ics.keplerian_disc(n, seed=n)
        o = O(p); o.record_neighbours(True); o.upload(b, s); o.evaluate()
        with E(p) as e:
            e.upload(b, s); e.evaluate()
            compare_eval(o, e)
    b, s = ics.keplerian_disc(1, seed=1)
    with E(p) as e:
        with pytest.raises(SphError):
            e.upload(b, s)


def test_run_until_matches_stepping(E):
    """sph_run_until (device-resident loop) == repeated sph_step."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(5_000, seed=8)
    with E(p) as e1, E(p) as e2:
        e1.upload(b, s); e2.upload(b, s)
        dt, t, steps = e1.run_until(0.05, 0.01, 0.0)
        dt2, t2, n2 = 0.01, 0.0, 0
        while t2 < 0.05:
            dt2, t2 = e2.step(dt2, t2); n2 += 1
        assert (dt, t, steps) == (dt2, t2, n2)
        b1, _ = e1.download(); b2, _ = e2.download()
        for k in GAS_FIELDS:
            assert np.array_equal(getattr(b1, k), getattr(b2, k)), k       # deterministic kernels


def test_million_particle_properties(E):
    """Size-independent properties at 1M particles (beyond the oracle's comfortable range): keys sorted,
    every particle is its own neighbour, SPH pair forces conserve momentum, downloads come back in
    ascending number order, repeat evaluations are bit-identical."""
    p = default_params(MODE_VARIABLE_H)
    n = 1_000_000
    b, s = ics.keplerian_disc(n, seed=5)
    with E(p) as e:
        e.upload(b, s)
        e.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_SPH)
        t = e.tree()
        assert np.array_equal(np.sort(t["order"]), np.arange(n))
        sorted_keys = t["key"][t["order"]]                                # keys along the DFS leaf order
        assert np.all(sorted_keys[1:] > sorted_keys[:-1])
        d = e.diag()
        assert np.all(d["rho"] >= b.m / (float(np.float32(np.pi)) * b.h ** 3) * (1 - 1e-12))   # self term
        for k in ("ax", "ay", "az"):
            tot = np.sum(b.m * d[k]); scale = np.sum(np.abs(b.m * d[k]))
            assert abs(tot) < 1e-11 * scale
        bb, _ = e.download()
        assert np.array_equal(bb.x, b.x) and np.array_equal(bb.h, b.h)
        e.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_SPH)
        d2 = e.diag()
        for k in ("rho", "ax", "udot"):
            assert np.array_equal(d[k], d2[k]), k


def test_disc_200k_vs_threaded_oracle(E, O):
    """200k disc against the oracle with OpenMP on its race-free loops (density, gravity, h iteration);
    the SPH pair loop stays serial in the oracle's parity path only when threads == 1, so compare the
    density / gravity phases here."""
    import os
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(200_000, seed=6)
    o = O(p, threads=min(8, os.cpu_count() or 1)); o.upload(b, s)
    mask = EVAL_TREE | EVAL_DENSITY | EVAL_GRAVITY | EVAL_SINKS
    o.evaluate(mask)
    with E(p) as e:
        e.upload(b, s); e.evaluate(mask)
        compare_eval(o, e, check_ngb=False, mask=mask)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_tree_reuse_is_bit_identical(mode, E, monkeypatch):
    """Evaluation A of a step sees the positions of the previous step's evaluation B (F:894 after F:905): the
    engine keeps that tree and only refreshes the reach R = 2h + size/2 with the new h (V:1152).  The state after
    several steps must equal, bit for bit, the state of a context that rebuilds the tree in every evaluation
    (which is what the oracle-checked tests above establish as the reference's result).  The far-field reuse of the
    gravity walk rides on the kept tree and changes the summation order, so it is off in both runs here; it has its
    own tests below."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(20_000, seed=5)
    out = []
    monkeypatch.setenv("SPH_B200_NO_FAR_REUSE", "1")
    for no_reuse in (False, True):
        if no_reuse:
            monkeypatch.setenv("SPH_B200_NO_TREE_REUSE", "1")
        else:
            monkeypatch.delenv("SPH_B200_NO_TREE_REUSE", raising=False)
        with E(p) as e:
            e.upload(b, s)
            dt, t = 0.01, 0.0
            for _ in range(4):
                dt, t = e.step(dt, t)
            be, se = e.download()
            out.append((dt, t, be, se, e.stage_times()))
    (dt0, t0, b0, s0, st0), (dt1, t1, b1, s1, st1) = out
    assert (dt0, t0) == (dt1, t1)
    for k in GAS_FIELDS:
        assert np.array_equal(getattr(b0, k), getattr(b1, k)), k
    for k in ("x", "y", "z", "vx", "vy", "vz", "m"):
        assert np.array_equal(getattr(s0, k), getattr(s1, k)), k


FAR_ENVS = {"lists": {}, "walk": {"SPH_B200_NO_FAR_LISTS": "1"}, "overflow": {"SPH_B200_FAR_SLOTS": "48"}, "off": {"SPH_B200_NO_FAR_REUSE": "1"}}


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_far_field_reuse_matches_full_walk(mode, E, monkeypatch):
    """Evaluation A of a loop body keeps the far-field gravity sums of the previous body's evaluation B (same positions,
    tree, sinks; F:894 after F:905-912) and adds only the near field with the new h - from the near pairs the full walk
    recorded ("lists"), by the near-only walk ("walk"), or both where some run's pairs did not fit their slots
    ("overflow").  Same terms as the full walk in another order: dt / t equal, state within rounding of a run that walks
    the whole tree in every evaluation ("off")."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(20_000, seed=5)
    out = {}
    for name, env in FAR_ENVS.items():
        for k in ("SPH_B200_NO_FAR_LISTS", "SPH_B200_FAR_SLOTS", "SPH_B200_NO_FAR_REUSE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with E(p) as e:
            e.upload(b, s)
            dt, t = 0.01, 0.0
            for _ in range(4):
                dt, t = e.step(dt, t)
            e.evaluate()
            out[name] = (dt, t, e.download(), e.diag(), e.counters(), e.far_reuse_count())
    dt1, t1, (b1, s1), d1, c1, n1 = out["off"]
    assert n1 == 0
    for name in ("lists", "walk", "overflow"):
        dt0, t0, (b0, s0), d0, c0, n0 = out[name]
        assert n0 >= 1, name
        if len(b0) == len(b):                 # nothing removed: evaluation B of steps 2-4 stores (the state stayed on the device since the
            assert n0 == 3, name              # step before), evaluation A of steps 3-4 and the evaluation after the last step reuse
        assert (dt0, t0) == (dt1, t1), name
        for k in GAS_FIELDS:
            assert relerr(getattr(b0, k), getattr(b1, k)) < 1e-12, (name, k)
        for k in ("x", "y", "z", "vx", "vy", "vz", "m"):
            assert relerr(getattr(s0, k), getattr(s1, k)) < 1e-12, (name, k)
        for k in d0:
            assert relerr(d0[k], d1[k]) < 1e-12, (name, k)
        assert c0 == c1, name                 # the near-only evaluation reports the counts of the accepted sets it stands for


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_far_field_reuse_vs_oracle(mode, E, O):
    """The near-only evaluation against the oracle's full evaluation of the same state (two sinks, accretion on: steps
    that remove particles void the stored sums and walk the whole tree)."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(10_000, seed=31)
    with E(p) as e:
        e.upload(b, s)
        dt, t = 0.01, 0.0
        for _ in range(3):
            dt, t = e.step(dt, t)
        be, se = e.download()
        n_before = e.far_reuse_count()
        e.evaluate()
        reused = e.far_reuse_count() - n_before
        o = O(p); o.record_neighbours(True); o.upload(be, se); o.evaluate()
        do, de = o.diag(), e.diag()
        for k in do:
            assert relerr(de[k], do[k]) < TOL, f"{k}: {relerr(de[k], do[k]):.3e}"
        assert o.counters()["grav_accepted"] == e.counters()["grav_accepted"]
        # a step that removed nothing leaves the sums standing
        if e.sizes() == (len(b), len(s)):
            assert reused == 1


def test_far_field_reuse_with_moving_sinks(E, O, monkeypatch):
    """Three sinks off the origin with masses that are not powers of two, moving, tiny accretion radii (nothing is
    removed): `initiate_sink_accretion` rewrites every sink position as (m x + 0) / m in every pass (F:497-501), so the
    sinks change from step to step.  The stored far sums hold tree terms only - the sink terms are taken anew in every
    evaluation - so the reuse stands, and the run follows the oracle."""
    p = default_params(MODE_VARIABLE_H)
    b, _ = ics.keplerian_disc(10_000, seed=23)
    s = Sinks([0.3, 41.0, -37.0], [-0.2, 3.0, 11.0], [0.01, 0.4, -0.3], [0.01, -0.4, 0.9], [0.02, 5.9, -5.1], [0.0, 0.01, 0.02],
              [0.7, 0.013, 0.021], [0.05, 0.05, 0.05])
    o = O(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        dto = dte = 0.01; to = te = 0.0
        for _ in range(4):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert dto == dte and to == te and o.sizes() == e.sizes()
        assert e.sizes() == (len(b), 3)
        assert e.far_reuse_count() == 2            # evaluation A of steps 3 and 4
        compare_state(o, e)


def test_far_field_reuse_h_cutoff(E, monkeypatch):
    """A smoothing length that grows beyond the cutoff its near / far split was taken with voids the stored sums: the
    next evaluation walks the whole tree.  Forced here with a cutoff of 1.000001 h (test hook; default 1.1 h): every
    step's calc_smoothing moves some h by more than that."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(10_000, seed=4)
    out = []
    for hcut in ("1.000001", None):
        if hcut:
            monkeypatch.setenv("SPH_B200_FAR_HCUT", hcut)
        else:
            monkeypatch.delenv("SPH_B200_FAR_HCUT", raising=False)
        with E(p) as e:
            e.upload(b, s)
            dt, t = 0.01, 0.0
            for _ in range(4):
                dt, t = e.step(dt, t)
            out.append((dt, t, e.download()[0], e.far_reuse_count()))
    (dt0, t0, b0, n0), (dt1, t1, b1, n1) = out
    assert n0 == 0 and n1 == 2                # steps 3 and 4 reuse unless the cutoff was exceeded
    assert (dt0, t0) == (dt1, t1)
    for k in GAS_FIELDS:
        assert relerr(getattr(b0, k), getattr(b1, k)) < 1e-12, k


@pytest.mark.parametrize("mode,removals", [(MODE_VARIABLE_H, False), (MODE_FIXED_H, False), (MODE_VARIABLE_H, True)])
def test_step_host_equals_upload_step_download(mode, removals, E):
    """sph_step_host (copies under the compute: late columns re-ordered when their first reader is due, early columns
    leaving mid-step) against sph_upload + sph_step + sph_download on the same rows: bit-identical, also when the step
    removes particles (the early columns are then sent again, compacted) and with the outputs aliasing the inputs."""
    p = default_params(mode) if not removals else default_params(mode, bounding_size=95.0)
    b, s = ics.keplerian_disc(20_000, seed=9)
    if removals:
        s.radius[:] = 12.0
    with E(p) as e1, E(p) as e2:
        b1, s1, b2, s2 = b, s, b.copy(), s
        dt1 = dt2 = 0.01; t1 = t2 = 0.0
        for k in range(3):
            e1.upload(b1, s1); dt1, t1 = e1.step(dt1, t1); b1, s1 = e1.download()
            ob = b2 if k == 1 else Bodies.empty(len(b2))              # step 1 writes into its own input arrays
            os_ = Sinks.empty(len(s2) + 8)
            dt2, t2, n2, ns2 = e2.step_host(b2, s2, dt2, t2, into=(ob, os_))
            b2 = Bodies(*[getattr(ob, f)[:n2].copy() for f in GAS_FIELDS]); s2 = Sinks(*[getattr(os_, f)[:ns2].copy() for f in ("x", "y", "z", "vx", "vy", "vz", "m", "radius")])
            assert (dt1, t1) == (dt2, t2) and (len(b1), len(s1)) == (n2, ns2)
            for f in GAS_FIELDS:
                assert np.array_equal(getattr(b1, f), getattr(b2, f)), (k, f)
            for f in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
                assert np.array_equal(getattr(s1, f), getattr(s2, f)), (k, f)
        if removals:
            assert len(b1) < len(b)


def test_step_host_two_word_keys(E):
    """Pairs closer than root_size/2^21: the first tree build of the call finds a 63-bit key collision and repeats itself
    with two-word keys - while the late columns of sph_step_host are still on their way (they are re-ordered before the
    second build).  Bit-identical to the three separate calls."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(5_000, seed=31)
    for a_, c_, eps in ((10, 11, 3e-7), (200, 201, 5e-9), (3000, 3001, 2e-10)):
        b.x[c_] = b.x[a_] + eps; b.y[c_] = b.y[a_] - 0.5 * eps; b.z[c_] = b.z[a_] + 0.25 * eps
    with E(p) as e1, E(p) as e2:
        e1.upload(b, s); e1.evaluate()
        assert np.max(e1.tree()["level"]) > 21                     # these rows do need the second key word
        e1.upload(b, s); dt1, t1 = e1.step(0.01, 0.0); b1, s1 = e1.download()
        ob, os_ = Bodies.empty(len(b)), Sinks.empty(len(s) + 8)
        dt2, t2, n2, ns2 = e2.step_host(b, s, 0.01, 0.0, into=(ob, os_))      # a fresh context: the retry happens inside this call
        assert (dt1, t1, len(b1), len(s1)) == (dt2, t2, n2, ns2)
        for f in GAS_FIELDS:
            assert np.array_equal(getattr(b1, f), getattr(ob, f)[:n2]), f


@pytest.mark.parametrize("mode", [MODE_VARIABLE_H, MODE_FIXED_H])
def test_resident_upload_is_recognised(mode, E):
    """A host that passes the state through every step (upload, step, download) hands back what it was given: the context
    recognises it (bitwise comparison of x y z m h + sinks on the device), keeps its tree and stored far-field sums and
    ends up - bit for bit - where a context that was never re-uploaded ends up.  Both host flows: three calls, one call."""
    p = default_params(mode)
    b, s = ics.keplerian_disc(20_000, seed=15)
    steps = 4
    with E(p) as e0:
        e0.upload(b, s)
        dt, t = 0.01, 0.0
        for _ in range(steps):
            dt, t = e0.step(dt, t)
        ref_b, ref_s = e0.download(); ref = (dt, t, e0.far_reuse_count())
    assert len(ref_b) == len(b)
    for flow in ("three calls", "one call"):
        with E(p) as e:
            bb, ss = b.copy(), s
            dt, t = 0.01, 0.0
            for k in range(steps):
                if flow == "three calls":
                    e.upload(bb, ss); dt, t = e.step(dt, t); bb, ss = e.download()
                else:
                    ob, os_ = Bodies.empty(len(bb)), Sinks.empty(len(ss) + 8)
                    dt, t, n2, ns2 = e.step_host(bb, ss, dt, t, into=(ob, os_))
                    bb = ob; ss = Sinks(*[getattr(os_, f)[:ns2].copy() for f in ("x", "y", "z", "vx", "vy", "vz", "m", "radius")])
            assert e.resident_hits() == steps - 1, flow
            assert (dt, t, e.far_reuse_count()) == ref, flow
            for f in GAS_FIELDS:
                assert np.array_equal(getattr(bb, f), getattr(ref_b, f)), (flow, f)
            for f in ("x", "y", "z", "vx", "vy", "vz", "m"):
                assert np.array_equal(getattr(ss, f), getattr(ref_s, f)), (flow, f)


def test_resident_upload_changed_state(E):
    """Same geometry, other velocities: recognised, the new v u alpha are taken over (result within rounding of a cold
    upload of the same rows, which walks the whole tree).  One position changed by one ulp: the comparison fails and the
    upload is a new state (bit-identical to the cold upload)."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(20_000, seed=16)
    with E(p) as e:
        e.upload(b, s)
        dt, t = 0.01, 0.0
        for _ in range(2):
            dt, t = e.step(dt, t)
        b1, s1 = e.download()                   # the context holds this state, out of a step
        for change in ("velocity", "position"):
            b2 = b1.copy()
            if change == "velocity":
                b2.vx[::7] *= 1.01; b2.u[::5] *= 0.99
            else:
                b2.x[123] = np.nextafter(b2.x[123], np.inf)
            h0 = e.resident_hits()
            e.upload(b2, s1); dta, ta = e.step(dt, t); ba, sa = e.download()
            assert e.resident_hits() - h0 == (1 if change == "velocity" else 0), change
            with E(p) as cold:
                cold.set_resident_check(False)
                cold.upload(b2, s1); dtc, tc = cold.step(dt, t); bc, sc_ = cold.download()
            assert (dta, ta) == (dtc, tc), change
            for f in GAS_FIELDS:
                if change == "velocity":
                    assert relerr(getattr(ba, f), getattr(bc, f)) < 1e-12, (change, f)
                else:
                    assert np.array_equal(getattr(ba, f), getattr(bc, f)), (change, f)
            b1, s1, dt, t = ba, sa, dta, ta     # again a state the context holds, out of a step


def test_candidate_list_pool_overflow_falls_back(E, monkeypatch):
    """The density pass saves the pair loop's candidate lists in a pool of blocks.  A pool that is too small voids
    the lists on the device: the walking pair kernel runs instead and the host doubles the pool for the next
    step.  Either way each target adds the same terms in the same order, so the states are bit-identical."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(30_000, seed=21)
    out = []
    for tiny in (False, True):
        if tiny:
            monkeypatch.setenv("SPH_B200_LIST_POOL_BLOCKS", "64")
        else:
            monkeypatch.delenv("SPH_B200_LIST_POOL_BLOCKS", raising=False)
        with E(p) as e:
            e.upload(b, s)
            dt, t = 0.01, 0.0
            for _ in range(3):
                dt, t = e.step(dt, t)
            e.evaluate()
            out.append((dt, t, e.download()[0], e.diag()))
    (dt0, t0, b0, d0), (dt1, t1, b1, d1) = out
    assert (dt0, t0) == (dt1, t1)
    for k in GAS_FIELDS:
        assert np.array_equal(getattr(b0, k), getattr(b1, k)), k
    for k in d0:
        assert np.array_equal(d0[k], d1[k]), k
