"""Profiling target: upload a disc of N particles, one warm-up evaluation, one profiled evaluation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summersph_b200 import default_params, MODE_VARIABLE_H, MODE_FIXED_H, ics
from summersph_b200.engine import Engine
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
mode = MODE_FIXED_H if (len(sys.argv) > 2 and sys.argv[2] == "F") else MODE_VARIABLE_H
p = default_params(mode)
b, s = ics.keplerian_disc(n)
with Engine(p) as e:
    e.upload(b, s)
    e.evaluate(); e.evaluate()
    print(e.stage_times(), e.counters())
