"""Native host-parallel text reader / writer (include/sph_textio.h) against the Python statement of the reference's
formats (summersph_b200/io.py, SUMMER_SPH.f90:594-738 | Variable.f90:729-942): identical rows, bit-identical values,
byte-identical save files, the reference's error cases.  CPU only."""
import ctypes as C
import os
import re
import time
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, ics, Bodies, Sinks
from summersph_b200 import io as pyio
from summersph_b200 import textio
from summersph_b200.state import GAS_FIELDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    textio.build()


def same(b0, s0, b1, s1):
    assert len(b0) == len(b1) and len(s0) == len(s1)
    for k in GAS_FIELDS:
        assert np.array_equal(getattr(b0, k), getattr(b1, k)), k
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
        assert np.array_equal(getattr(s0, k), getattr(s1, k)), k


def test_library_exports_every_declared_symbol():
    txt = open(os.path.join(ROOT, "include", "sph_textio.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    lib = C.CDLL(textio.LIB_PATH)
    syms = sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", txt)))
    assert "sph_ics_open" in syms and "sph_save_write" in syms
    for s in syms:
        assert hasattr(lib, s), s


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_reader_matches_python_reader(tmp_path, mode, threads):
    p = default_params(mode)
    b, s = ics.keplerian_disc(5000, seed=3)
    fn = str(tmp_path / "disc.txt")
    pyio.write_ics(fn, b, s, columns=10)
    b0, s0 = pyio.read_data_from_file(fn, p)
    b1, s1 = textio.read_data_from_file(fn, p, threads=threads, log=None)
    same(b0, s0, b1, s1)
    if mode == MODE_FIXED_H:
        assert np.all(b1.alpha == 0.0) and np.all(b1.h == p.h_fixed)          # F:681
    assert np.all(s1.radius == p.sink_radius)                                  # F:694 | V:830


def test_reader_edge_cases(tmp_path):
    pv, pf = default_params(MODE_VARIABLE_H), default_params(MODE_FIXED_H)
    fn = str(tmp_path / "edge.txt")
    # sink rows anywhere with 8 columns, blank rows, extra columns, tabs / commas, Fortran D exponents, '+' signs,
    # no trailing newline
    open(fn, "w").write("x y z vx vy vz energy mass alpha smoothing\n"
                        "1.0 2.0 3.0 0.1 0.2 0.3 0.25 1e-6 0.1 2.5 99 98\n"
                        "\n"
                        "0 0 0 0 0 0 0.0 1.0\n"
                        "\t+1.5d0,2.0D+00 ,3 0.1 0.2 0.3 2.5E-1 1.0e-06 0.2 3.5\n"
                        "   \n"
                        "-1 -2 -3 0 0 0 0.5 2e-6 0.3 1.5")
    b, s = textio.read_data_from_file(fn, pv, log=None)
    assert len(b) == 3 and len(s) == 1
    assert b.x.tolist() == [1.0, 1.5, -1.0] and b.h.tolist() == [2.5, 3.5, 1.5] and b.alpha.tolist() == [0.1, 0.2, 0.3]
    assert s.m.tolist() == [1.0] and s.radius.tolist() == [pv.sink_radius]
    b, s = textio.read_data_from_file(fn, pf, log=None)                        # fixed h: 8 columns, alpha := 0
    assert len(b) == 3 and b.alpha.tolist() == [0.0, 0.0, 0.0] and b.h.tolist() == [pf.h_fixed] * 3
    # no sink row -> the dummy sink (F:698-707)
    open(fn, "w").write("hdr\n1 2 3 0 0 0 0.25 1e-6 0.1 2.5\n")
    b, s = textio.read_data_from_file(fn, pv, log=None)
    assert len(b) == 1 and len(s) == 1 and s.m[0] == 0.0 and s.radius[0] == 0.0
    # a short row / a non-number: "Error reading line N" with the data-row number (F:648-651)
    open(fn, "w").write("hdr\n1 2 3 0 0 0 0.25 1e-6 0.1 2.5\n1 2 3 0 0 0 0.25\n")
    with pytest.raises(ValueError, match="Error reading line 2"):
        textio.read_data_from_file(fn, pv, log=None)
    open(fn, "w").write("hdr\n1 2 3 0 0 0 0.25 1e-6 0.1\n")                    # gas row needs 10 columns in variable h
    with pytest.raises(ValueError, match="Error reading line 1"):
        textio.read_data_from_file(fn, pv, log=None)
    textio.read_data_from_file(fn, pf, log=None)                               # ... but only 8 in fixed h
    open(fn, "w").write("hdr\n1 2 x 0 0 0 0.25 1e-6 0.1 2.5\n")
    with pytest.raises(ValueError, match="Error reading line 1"):
        textio.read_data_from_file(fn, pv, log=None)
    open(fn, "w").write("only a header\n")
    with pytest.raises(ValueError, match="No data found"):                     # F:625-628
        textio.read_data_from_file(fn, pv, log=None)
    with pytest.raises(FileNotFoundError, match="Error opening file"):         # F:612-615
        textio.read_data_from_file(str(tmp_path / "missing.txt"), pv, log=None)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_save_is_byte_identical_and_round_trips(tmp_path, mode):
    p = default_params(mode)
    b, s = ics.keplerian_disc(3000, seed=8)
    b.x[5] = 1.2345678901234567e-300; b.y[6] = -9.87654321e+200; b.u[7] = 1e-310     # three-digit exponents, a subnormal
    d0, d1 = tmp_path / "py", tmp_path / "native"
    d0.mkdir(); d1.mkdir()
    f0 = pyio.make_save(b, s, 4, p, str(d0))
    f1 = textio.make_save(b, s, 4, p, str(d1), threads=3)
    assert open(f0, "rb").read() == open(f1, "rb").read()
    with pytest.raises(FileExistsError):                                       # status="new" (F:728)
        textio.make_save(b, s, 4, p, str(d1))
    # a save file is a valid IC file (resume): 17 significant digits are lossless
    pr = default_params(MODE_VARIABLE_H) if mode == MODE_VARIABLE_H else p
    b1, s1 = textio.read_data_from_file(f1, pr, log=None)
    for k in ("x", "y", "z", "vx", "vy", "vz", "u", "m"):
        assert np.array_equal(getattr(b1, k), getattr(b, k)), k
    if mode == MODE_VARIABLE_H:
        assert np.array_equal(b1.alpha, b.alpha) and np.array_equal(b1.h, b.h)
    assert np.array_equal(s1.m, s.m)


def test_throughput_against_the_python_reader(tmp_path):
    """Measurement beside parity: the same 200k-row file through both readers and both writers."""
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(200_000, seed=1)
    d0, d1 = tmp_path / "py", tmp_path / "native"
    d0.mkdir(); d1.mkdir()
    t0 = time.perf_counter(); f0 = pyio.make_save(b, s, 0, p, str(d0)); t_wpy = time.perf_counter() - t0
    t0 = time.perf_counter(); f1 = textio.make_save(b, s, 0, p, str(d1)); t_wna = time.perf_counter() - t0
    mb = os.path.getsize(f1) / 1e6
    t0 = time.perf_counter(); b0, s0 = pyio.read_data_from_file(f0, p); t_rpy = time.perf_counter() - t0
    t0 = time.perf_counter(); b1, s1 = textio.read_data_from_file(f1, p, log=None); t_rna = time.perf_counter() - t0
    same(b0, s0, b1, s1)
    print(f"\n{mb:.0f} MB save file: write {mb / t_wpy:.0f} -> {mb / t_wna:.0f} MB/s, read {mb / t_rpy:.0f} -> {mb / t_rna:.0f} MB/s "
          f"(python -> native, {os.cpu_count()} cores)")
    assert t_rna < t_rpy and t_wna < t_wpy
