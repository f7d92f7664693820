"""Text surface of the reference: IC reader, parameters.txt, save files, simulate shell."""
import os
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, ics, Bodies, Sinks
from summersph_b200.io import read_data_from_file, read_params_from_file, make_save, write_ics
from summersph_b200.simulate import simulate


@pytest.fixture
def disc():
    b, s = ics.keplerian_disc(50, seed=5)
    s.radius[:] = 5.0
    return b, s


def test_reader_fixed_h_semantics(tmp_path, disc):
    b, s = disc
    f = tmp_path / "disc_12000_2.txt"
    write_ics(f, b, s, columns=9)                       # Disc_ICs.py writes 9 columns + header
    p = default_params(MODE_FIXED_H)
    b2, s2 = read_data_from_file(str(f), p)
    assert len(b2) == 50 and len(s2) == 1
    assert np.all(b2.alpha == 0.0)                      # F:681
    assert np.all(b2.h == 2.5) and s2.radius[0] == 3.5  # F:11, F:694
    assert s2.m[0] == 1.0
    assert np.allclose(b2.x, b.x, rtol=1e-15)


def test_reader_variable_h_and_sink_rows(tmp_path, disc):
    b, s = disc
    f = tmp_path / "disc_20k_low_vel.txt"
    write_ics(f, b, s, columns=10)
    p = default_params(MODE_VARIABLE_H)
    b2, s2 = read_data_from_file(str(f), p)
    assert np.allclose(b2.h, b.h, rtol=1e-15) and np.allclose(b2.alpha, 0.1)
    assert s2.radius[0] == 5.0                          # V:830
    # the header line is always skipped (a file without one loses its first particle, F:617)
    lines = open(f).read().splitlines()
    (tmp_path / "nohdr.txt").write_text("\n".join(lines[1:]) + "\n")
    b3, _ = read_data_from_file(str(tmp_path / "nohdr.txt"), p)
    assert len(b3) == 49
    # u == 0 marks a sink wherever it appears; gas numbering = order of appearance among gas rows (F:684)
    rows = lines[1:]
    mixed = [lines[0], rows[0], rows[-1]] + rows[1:-1]
    (tmp_path / "mixed.txt").write_text("\n".join(mixed) + "\n")
    b4, s4 = read_data_from_file(str(tmp_path / "mixed.txt"), p)
    assert len(s4) == 1 and np.allclose(b4.x, b.x, rtol=1e-15)


def test_reader_no_sink_creates_dummy(tmp_path, disc):
    b, _ = disc
    f = tmp_path / "ic.txt"
    write_ics(f, b, None)
    _, s = read_data_from_file(str(f), default_params(MODE_VARIABLE_H))
    assert len(s) == 1 and s.m[0] == 0.0 and s.radius[0] == 0.0      # F:698-707


def test_reader_errors(tmp_path):
    p = default_params(MODE_FIXED_H)
    with pytest.raises(FileNotFoundError):
        read_data_from_file(str(tmp_path / "missing.txt"), p)
    (tmp_path / "empty.txt").write_text("header only\n")
    with pytest.raises(ValueError):
        read_data_from_file(str(tmp_path / "empty.txt"), p)
    (tmp_path / "short.txt").write_text("h\n1 2 3 4 5 6 7\n")
    with pytest.raises(ValueError):
        read_data_from_file(str(tmp_path / "short.txt"), p)


def test_params_file(tmp_path):
    f = tmp_path / "parameters.txt"
    f.write_text("bounding max_depth theta gamma eta conv max_len scale end\n1000 30 0.7 1.6666 1.3 0.01 40 0.3 2.5\n900 21 0.5 1.4 1.2 0.001 50 0.25 0.1\n")
    p = read_params_from_file(str(f))
    assert (p.bounding_size, p.max_depth, p.theta, p.gamma, p.eta) == (900.0, 21, 0.5, 1.4, 1.2)   # last row wins
    assert (p.convergence_criteria, p.max_length, p.timestep_scale, p.end_time) == (0.001, 50.0, 0.25, 0.1)
    assert p.mode & MODE_VARIABLE_H


def test_save_roundtrip_and_status_new(tmp_path, disc):
    b, s = disc
    p = default_params(MODE_VARIABLE_H)
    path = make_save(b, s, 3, p, str(tmp_path))
    assert os.path.basename(path) == "save3.txt"
    with pytest.raises(FileExistsError):                # status="new", F:728
        make_save(b, s, 3, p, str(tmp_path))
    b2, s2 = read_data_from_file(path, p)               # V re-reads its own saves (sink rows last)
    for k in ("x", "y", "z", "vx", "vy", "vz", "u", "m", "alpha", "h"):
        assert np.array_equal(getattr(b2, k), getattr(b, k)), k
    assert np.array_equal(s2.m, s.m)
    # fixed-h saves have 9 gas columns and re-read with alpha reset to 0
    pf = default_params(MODE_FIXED_H)
    path = make_save(b, s, 0, pf, str(tmp_path))
    assert len(open(path).read().splitlines()[1].split()) == 9
    b3, _ = read_data_from_file(path, pf)
    assert np.all(b3.alpha == 0.0) and np.array_equal(b3.u, b.u)


class FakeEngine:
    """Host-logic stand-in: advances t by dt, grows dt by 1.5 (no physics)."""
    def __init__(self):
        self.calls = 0
    def upload(self, b, s):
        self.b, self.s = b, s
    def sizes(self):
        return len(self.b), len(self.s)
    def step(self, dt, t):
        self.calls += 1
        return min(dt * 1.5, 0.09), t + dt
    def download(self):
        return self.b, self.s


def test_simulate_shell_end_time_and_saves(tmp_path, disc):
    b, s = disc
    p = default_params(MODE_VARIABLE_H, end_time=0.1)
    lines = []
    eng = FakeEngine()
    _, _, t, dt, steps = simulate(b, s, p, engine=eng, save_dir=str(tmp_path), log=lines.append)
    assert t >= 0.1 and steps == eng.calls               # first step with t >= end_time stops the loop (F:879)
    assert lines[0].startswith(" SPH Particles: 50 dt : 0.01 time :  0.0")   # F:891
    saves = sorted(int(f[4:-4]) for f in os.listdir(tmp_path) if f.startswith("save"))
    assert saves == list(range(len(saves))) and len(saves) == steps - 1      # one save per pass after the first
