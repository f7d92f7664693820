"""Sink spin bookkeeping and the sink merger — opt-in with FLAG_SINK_MERGE_SPIN.

Not the reference's behaviour: `sink%spin` is declared (F:33), zeroed (F:695, V:580) and never updated ("also need
something to track the angular momentum", F:509), and `check_sink_merger` is an empty stub whose call is commented out
(V:1067-1073, V:1159).  The definitions are this build's own (include/sph_b200.h, oracle/sph_oracle.cpp):
  accretion : spin += L_orbital(sink + accreted gas) before - L_orbital(sink) after the reference's mass-weighted merge
  merger    : |x_a - x_b| < max(R_a, R_b), both with mass -> lower index keeps summed mass, mass-weighted x v a,
              larger radius, spin = S_a + S_b + L_orbital before - after; higher index removed, order preserved.
Without the flag nothing changes (golden fixtures and every parity test run with it off).
CPU tests pin the oracle's statement by conservation laws; the GPU tests compare the CUDA kernels with it."""
import numpy as np
import pytest

from summersph_b200 import (default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SINK_MERGE_SPIN, ics, Bodies, Sinks)
from oracle.oracle import Oracle

MODE = MODE_VARIABLE_H | FLAG_SINK_MERGE_SPIN


def orbital_L(m, x, y, z, vx, vy, vz):
    return np.array([np.sum(m * (y * vz - z * vy)), np.sum(m * (z * vx - x * vz)), np.sum(m * (x * vy - y * vx))])


def sinks_L(s):
    return orbital_L(s.m, s.x, s.y, s.z, s.vx, s.vy, s.vz)


def three_sinks():
    """0 and 2 overlap (|dx| = 3 < R_0 = 4), 1 is far away; all carry mass and move."""
    return Sinks([10.0, -60.0, 13.0], [5.0, 0.0, 5.0], [0.0, 1.0, 0.5], [0.3, 0.0, -0.2], [1.0, -2.0, 1.4], [0.0, 0.1, 0.05],
                 [1.0, 0.5, 0.25], [4.0, 2.0, 1.0])


def few_gas(n=50, seed=1):
    b, _ = ics.keplerian_disc(n, seed=seed)
    return b


def test_merger_conserves_mass_momentum_and_total_angular_momentum():
    o = Oracle(default_params(MODE)); s = three_sinks(); o.upload(few_gas(), s)
    P0 = np.array([np.sum(s.m * s.vx), np.sum(s.m * s.vy), np.sum(s.m * s.vz)]); L0 = sinks_L(s)
    o.check_sink_merger()
    _, s1 = o.download(); spin = o.sink_spin()
    assert len(s1) == 2 and spin.shape == (2, 3)
    assert s1.m[0] == 1.25 and s1.m[1] == 0.5 and s1.radius[0] == 4.0 and s1.radius[1] == 2.0          # order preserved
    assert s1.x[0] == pytest.approx((10.0 * 1.0 + 13.0 * 0.25) / 1.25) and s1.vy[0] == pytest.approx((1.0 + 0.25 * 1.4) / 1.25)
    P1 = np.array([np.sum(s1.m * s1.vx), np.sum(s1.m * s1.vy), np.sum(s1.m * s1.vz)])
    assert np.allclose(P1, P0, rtol=0, atol=1e-15)
    assert np.allclose(sinks_L(s1) + spin.sum(0), L0, rtol=0, atol=1e-13)
    assert np.all(spin[1] == 0.0) and np.linalg.norm(spin[0]) > 1e-3    # relative motion of the pair became spin
    o.check_sink_merger()                                               # nothing left to merge
    assert o.sizes()[1] == 2


def test_merger_chain_and_massless_sinks():
    """a-b and (a+b)-c overlap only after the first merge: the scan restarts; a zero-mass (dummy) sink never merges."""
    s = Sinks([0.0, 1.5, 4.0, 0.2], [0.0] * 4, [0.0] * 4, [0.0] * 4, [0.0, 1.0, -1.0, 0.0], [0.0] * 4,
              [1.0, 1.0, 1.0, 0.0], [2.0, 1.0, 3.4, 9.0])
    o = Oracle(default_params(MODE)); o.upload(few_gas(), s)
    o.check_sink_merger()
    _, s1 = o.download()
    # 0+1 (|dx| 1.5 < 2) -> x = 0.75; then with 2: |4 - 0.75| = 3.25 < 3.4 -> one sink of mass 3 at x = 11/6; the massless one stays
    assert len(s1) == 2 and s1.m[0] == 3.0 and s1.m[1] == 0.0 and s1.x[0] == pytest.approx(5.5 / 3.0) and s1.radius[0] == 3.4
    assert np.allclose(sinks_L(s1) + o.sink_spin().sum(0), sinks_L(s), atol=1e-14)


def test_accretion_moves_orbital_angular_momentum_into_spin():
    """The reference's mass-weighted merge (F:497-508) conserves mass and momentum but not sum m x cross v; with the
    flag the difference is kept as spin, without it the spin stays zero like the reference's."""
    b, _ = ics.keplerian_disc(3000, seed=9)
    s = Sinks([0.0, 40.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 6.0], [0.0, 0.0], [1.0, 0.01], [20.0, 6.0])
    L_gas = orbital_L(b.m, b.x, b.y, b.z, b.vx, b.vy, b.vz)
    out = {}
    for flag in (0, FLAG_SINK_MERGE_SPIN):
        o = Oracle(default_params(MODE_VARIABLE_H | flag)); o.upload(b, s)
        o.evaluate(1)                        # tree only (sink2gasdists walks it, V:649-676)
        o.accrete()
        b1, s1 = o.download()
        assert len(b1) < len(b)
        out[flag] = (orbital_L(b1.m, b1.x, b1.y, b1.z, b1.vx, b1.vy, b1.vz) + sinks_L(s1), o.sink_spin(), s1)
    (L_off, spin_off, s_off), (L_on, spin_on, s_on) = out[0], out[FLAG_SINK_MERGE_SPIN]
    assert np.all(spin_off == 0.0)
    for k in ("x", "y", "z", "vx", "vy", "vz", "m"):                    # the flag never changes the reference's merge itself
        assert np.array_equal(getattr(s_off, k), getattr(s_on, k)), k
    L0 = L_gas + sinks_L(s)
    assert np.linalg.norm(L_off - L0) > 1e-6 * np.linalg.norm(L0)       # orbital L alone is not conserved by F:497-508
    assert np.allclose(L_on + spin_on.sum(0), L0, rtol=0, atol=1e-12 * np.linalg.norm(L0))
    assert np.linalg.norm(spin_on[0]) > 0 and np.linalg.norm(spin_on[1]) > 0


def test_flag_off_is_the_reference_and_flag_on_merges_in_step():
    """A step with two overlapping massive sinks: off -> both survive (the reference never merges); on -> one sink,
    and the conserved sums' L includes the spin."""
    b, _ = ics.keplerian_disc(1500, seed=4)
    s = Sinks([0.0, 2.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 3.0], [0.0, 0.0], [1.0, 0.05], [5.0, 5.0])
    n_sinks = {}
    for flag in (0, FLAG_SINK_MERGE_SPIN):
        o = Oracle(default_params(MODE_VARIABLE_H | flag)); o.upload(b, s)
        L0 = np.array([o.conserved()[k] for k in ("lx", "ly", "lz")])
        o.step(0.01, 0.0)
        n_sinks[flag] = o.sizes()[1]
        c = o.conserved()
        if flag:
            _, s1 = o.download(); b1, _ = o.download()
            Lsum = orbital_L(b1.m, b1.x, b1.y, b1.z, b1.vx, b1.vy, b1.vz) + sinks_L(s1) + o.sink_spin().sum(0)
            assert np.allclose([c["lx"], c["ly"], c["lz"]], Lsum, rtol=1e-12, atol=1e-15)
            assert abs(c["lz"] - L0[2]) < 1e-4 * abs(L0[2])            # one step of forces only
    assert n_sinks == {0: 2, FLAG_SINK_MERGE_SPIN: 1}


# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def E(built_engine):
    from summersph_b200.engine import Engine
    return Engine


def compare_sinks(e, o, L_scale):
    be, se = e.download(); bo, so = o.download()
    assert len(be) == len(bo) and len(se) == len(so)
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
        a, r = getattr(se, k), getattr(so, k)
        assert np.allclose(a, r, rtol=1e-10, atol=1e-12), k
    assert np.allclose(e.sink_spin(), o.sink_spin(), rtol=0, atol=1e-10 * L_scale)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_gpu_spin_and_merger_match_oracle(mode, E):
    """Accreting, overlapping sinks over three steps: sizes, sink state and spin follow the oracle."""
    p = default_params(mode | FLAG_SINK_MERGE_SPIN, bounding_size=85.0)
    b, _ = ics.keplerian_disc(8_000, seed=9)
    s = Sinks([0.0, 40.0, 43.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.2], [0.0, 6.0, 5.5], [0.0, 0.0, 0.0],
              [1.0, 0.01, 0.02], [13.0, 6.0, 2.0])
    L_scale = np.linalg.norm(orbital_L(b.m, b.x, b.y, b.z, b.vx, b.vy, b.vz)) + np.linalg.norm(sinks_L(s))
    o = Oracle(p); o.upload(b, s)
    with E(p) as e:
        e.upload(b, s)
        assert np.all(e.sink_spin() == 0.0)
        dto = dte = 0.01; to = te = 0.0
        for k in range(3):
            dto, to = o.step(dto, to); dte, te = e.step(dte, te)
            assert o.sizes() == e.sizes(), k
            assert (dto, to) == (dte, te)
            compare_sinks(e, o, L_scale)
            ce, co = e.conserved(), o.conserved()
            for q in ("lx", "ly", "lz"):
                assert abs(ce[q] - co[q]) <= 1e-10 * L_scale, q
        assert o.sizes()[1] == 2 and o.sizes()[0] < 8_000                # 1 and 2 merged, gas was accreted
        assert np.linalg.norm(o.sink_spin()) > 0


@pytest.mark.gpu
def test_gpu_flag_off_keeps_spin_zero_and_never_merges(E):
    p = default_params(MODE_VARIABLE_H, bounding_size=85.0)
    b, _ = ics.keplerian_disc(4_000, seed=9)
    s = Sinks([0.0, 2.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 3.0], [0.0, 0.0], [1.0, 0.05], [13.0, 5.0])
    with E(p) as e:
        e.upload(b, s)
        e.step(0.01, 0.0)
        assert e.sizes()[1] == 2 and np.all(e.sink_spin() == 0.0)
