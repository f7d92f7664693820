"""The hosts' drift report (`run_sph --drift`, `simulate(drift=...)`): the C++ twin of `program run_sph` and the Python
host print the same conserved sums for the same IC file (GPU; the CPU side of `simulate(drift=...)` is covered in
tests/test_widen_conserved.py)."""
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, ics


@pytest.mark.gpu
def test_cpp_host_drift_report_matches_python_host(tmp_path, built_engine):
    """host/run_sph --drift prints the sums before the first and after the last step; the Python host driving the same
    engine on the same IC file sees the same numbers."""
    import os, re, subprocess
    from summersph_b200.engine import Engine
    from summersph_b200.io import write_ics, read_data_from_file
    from summersph_b200.simulate import simulate
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "host"), "-s"])
    p = default_params(MODE_VARIABLE_H, end_time=0.04)
    b, s = ics.keplerian_disc(3000, seed=5)
    s.radius[:] = p.sink_radius
    ic = tmp_path / "ics.txt"
    write_ics(ic, b, s)
    out = subprocess.run([os.path.join(root, "host", "run_sph"), "--variable", "--end-time", "0.04", "--drift", str(ic)],
                         capture_output=True, text=True, check=True).stdout
    m = re.search(r"Conserved sums: E = (\S+) -> (\S+)\s+\(kin (\S+) int (\S+) pot (\S+)\)", out)
    d = re.search(r"Drift over (\d+) steps: dE/\|E0\| = (\S+)", out)
    assert m and d, out[-400:]
    bb, ss = read_data_from_file(str(ic), p)
    rep = {}
    with Engine(p) as e:
        simulate(bb, ss, p, engine=e, log=lambda *_: None, drift=rep)
    assert int(d.group(1)) == rep["steps"]
    for got, want in ((m.group(1), rep["first"]["e_total"]), (m.group(2), rep["last"]["e_total"]), (m.group(3), rep["last"]["e_kin"]),
                      (m.group(4), rep["last"]["e_int"]), (m.group(5), rep["last"]["e_pot"])):
        assert float(got) == pytest.approx(want, rel=1e-13)
    assert float(d.group(2)) == pytest.approx(rep["energy_rel"], rel=1e-3, abs=1e-12)
