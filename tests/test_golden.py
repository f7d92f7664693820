"""Golden fixtures (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py).

CPU: the oracle reproduces its committed outputs bit for bit (regression pin of the restatement).
GPU: the CUDA engine, called through the C-ABI, matches them: integer/order data bit-exact, FP64 fields
to 1e-10 relative (north_star tolerance; scale = max(|value|, field RMS))."""
import numpy as np
import pytest

from _cases import GOLDEN, load_golden, run_case
from conftest import relerr

TOL = 1e-10
EXACT = ("tree_order", "tree_level", "tree_cx", "tree_cy", "tree_cz", "tree_size", "ngb_count", "ngb_hash", "st_dt_t")


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_reproduces_golden(name):
    from oracle.oracle import Oracle
    z, p, b, s = load_golden(name)
    o = Oracle(p); o.record_neighbours(True)
    out = run_case(o, z, b, s)
    for k, v in out.items():
        assert v.shape == z[k].shape, k
        assert np.array_equal(v, z[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_engine_matches_golden(name, built_engine):
    from summersph_b200.engine import Engine
    z, p, b, s = load_golden(name)
    with Engine(p) as e:
        out = run_case(e, z, b, s)
    for k, v in out.items():
        assert v.shape == z[k].shape, f"{k}: {v.shape} vs {z[k].shape}"
        if k in EXACT:
            assert np.array_equal(v, z[k]), k
        else:
            assert relerr(v, z[k]) < TOL, f"{k}: {relerr(v, z[k]):.3e}"
