"""The C++ twin of `program run_sph` (host/run_sph, linked only against the C-ABI library) driven end to end
from an IC text file: its per-step log and its save files match the Python host driving the same engine."""
import os
import re
import subprocess
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, MODE_FIXED_H, ics
from summersph_b200.io import write_ics, read_data_from_file

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("variable", [True, False])
def test_run_sph_cpp_matches_python_host(variable, tmp_path, built_engine):
    from summersph_b200.engine import Engine
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    mode = MODE_VARIABLE_H if variable else MODE_FIXED_H
    p = default_params(mode, end_time=0.06)
    b, s = ics.keplerian_disc(4000, seed=77)
    s.radius[:] = p.sink_radius
    ic = tmp_path / "ics.txt"
    write_ics(ic, b, s)
    save_dir = tmp_path / "saves"; save_dir.mkdir()
    cmd = [os.path.join(ROOT, "host", "run_sph")] + (["--variable"] if variable else []) + ["--end-time", "0.06", "--save-dir", str(save_dir), str(ic)]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    log = [(int(m.group(1)), float(m.group(2)), float(m.group(3))) for m in re.finditer(r"SPH Particles:\s*(\d+)\s*dt :\s*(\S+)\s*time :\s*(\S+)", out)]
    assert "Successfully read 4000 bodies and 1 sinks" in out and len(log) >= 3

    bb, ss = read_data_from_file(str(ic), p)          # the same text file through the Python reader
    with Engine(p) as e:
        e.upload(bb, ss)
        t, dt, k = 0.0, 1e-2, 0
        while t < p.end_time:
            assert log[k] == (e.sizes()[0], dt, t)     # F:891 line by line
            dt, t = e.step(dt, t); k += 1
        assert k == len(log)
    saves = sorted(os.listdir(save_dir), key=lambda f: int(f[4:-4]))
    assert saves == [f"save{i}.txt" for i in range(len(log) - 1)]
    # save_k holds the state before step k+1, i.e. after k+1 steps from the ICs (save0 is written at the 2nd pass)
    with Engine(p) as e:
        e.upload(bb, ss)
        dt, t = e.step(1e-2, 0.0)
        eb, es = e.download()
    sb, sk = read_data_from_file(str(save_dir / "save0.txt"), p)
    for f in ("x", "y", "z", "vx", "vy", "vz", "u", "m") + (("alpha", "h") if variable else ()):
        assert np.array_equal(getattr(sb, f), getattr(eb, f)), f
    assert np.array_equal(sk.m, es.m) and np.array_equal(sk.x, es.x)
