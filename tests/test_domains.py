"""Morton-domain decomposition (params.decomposition = 1): N ranks, one process per rank, against the
single-rank run of the same case.

What must be IDENTICAL (integer / index work): the global Morton order, every particle's key, leaf level and cell, the
neighbour sets (counts + hashes), the interaction counters, the particle and sink counts, dt and t.
What must agree to 1e-12 relative per evaluation (1e-10 after three loop bodies): every FP64 field - the same terms are
summed, grouped differently only where a walk group or a gravity run meets a domain boundary.

The ranks share cuda:0 through the host-segment communicator (runs on the 1-GPU box); with >= 2 GPUs the NCCL form
runs as well."""
import numpy as np
import pytest

from summersph_b200 import MODE_VARIABLE_H, MODE_FIXED_H
from test_multi_gpu import run_ranks, _ngpu

pytestmark = pytest.mark.gpu

INT_KEYS = ("t_order", "t_key", "t_level", "t_cx", "t_cy", "t_cz", "t_size", "n_count", "n_hash", "e_counters")
FP_EVAL = ("e_rho", "e_omega", "e_P", "e_c", "e_ax", "e_ay", "e_az", "e_udot", "e_alphadot", "e_sink_ax", "e_sink_ay", "e_sink_az")
STATE = ("x", "y", "z", "vx", "vy", "vz", "u", "m", "alpha", "h", "s_x", "s_y", "s_z", "s_vx", "s_vy", "s_vz", "s_m")


def rel(a, b):
    scale = np.maximum(np.abs(b), np.sqrt(np.mean(b * b)) if b.size else 1.0)
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def compare(res, ref, what, tol_eval, tol_state):
    assert np.array_equal(res["meta"], ref["meta"]), (what, "dt / t / sizes", res["meta"], ref["meta"])
    for k in INT_KEYS:
        a, b = res[k], ref[k]
        if k == "e_counters":          # n_nodes (index 6 in sorted-key order: ..., n_gas, n_nodes, sph_pairs) is per rank under domains
            names = sorted(["n_gas", "n_nodes", "density_candidates", "density_contributing", "sph_pairs", "grav_opened", "grav_accepted", "h_iterations"])
            keep = [i for i, nm in enumerate(names) if nm not in ("n_nodes", "grav_opened")]
            a, b = a[keep], b[keep]
        assert a.shape == b.shape and np.array_equal(a, b), (what, k)
    worst = {k: rel(res[k], ref[k]) for k in FP_EVAL}
    assert max(worst.values()) < tol_eval, (what, worst)
    worst_s = {k: rel(res[k], ref[k]) for k in STATE}
    assert max(worst_s.values()) <= tol_state, (what, worst_s)
    return worst, worst_s


@pytest.fixture(scope="module")
def reference(tmp_path_factory, built_engine):
    d = tmp_path_factory.mktemp("dd_ref")
    out = {}
    for mode in (MODE_VARIABLE_H, MODE_FIXED_H):
        for steps in (0, 3):
            out[(mode, steps)] = run_ranks(d, f"ref{mode}_{steps}", 1, "host", mode, steps=steps, extra="tree")[0]
    return out


@pytest.mark.parametrize("mode", [MODE_VARIABLE_H, MODE_FIXED_H])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_domains_one_evaluation(mode, world, reference, tmp_path):
    res = run_ranks(tmp_path, f"dd{world}", world, "host", mode, steps=0, domains=1, extra="tree")
    for r in range(world):
        compare(res[r], reference[(mode, 0)], f"world {world} rank {r} eval", 1e-12, 0.0)
    assert sum(int(res[r]["local_n"][0]) for r in range(world)) == int(reference[(mode, 0)]["meta"][2])
    assert max(int(res[r]["local_n"][0]) for r in range(world)) < 1.05 * reference[(mode, 0)]["meta"][2] / world + 64      # balanced domains


@pytest.mark.parametrize("mode", [MODE_VARIABLE_H, MODE_FIXED_H])
@pytest.mark.parametrize("world", [2, 4])
def test_domains_three_steps(mode, world, reference, tmp_path):
    res = run_ranks(tmp_path, f"dds{world}", world, "host", mode, steps=3, domains=1, extra="tree")
    for r in range(world):
        compare(res[r], reference[(mode, 3)], f"world {world} rank {r} steps", 1e-10, 1e-10)


@pytest.mark.parametrize("world", [2, 3])
def test_domains_far_reuse(world, built_engine, tmp_path):
    """A case that removes nothing, four loop bodies: every domain keeps evaluation B's far-field gravity for the next
    evaluation A (the ranks agree on it through one all-reduce; the sinks' own sums are the stored ones)."""
    env = {"SPH_TEST_QUIET": "1"}
    ref = run_ranks(tmp_path, "ddq1", 1, "host", MODE_VARIABLE_H, steps=4, env_extra=env, n=30_000, extra="tree")[0]
    res = run_ranks(tmp_path, f"ddq{world}", world, "host", MODE_VARIABLE_H, steps=4, env_extra=env, n=30_000, domains=1, extra="tree")
    assert int(ref["far_reuse"][0]) >= 2
    for r in range(world):
        assert int(res[r]["far_reuse"][0]) == int(ref["far_reuse"][0]), (r, res[r]["far_reuse"], ref["far_reuse"])
        compare(res[r], ref, f"quiet world {world} rank {r}", 1e-10, 1e-10)


@pytest.mark.parametrize("world", [2, 3])
def test_domains_recognise_handed_back_state(world, built_engine, tmp_path):
    """Every rank downloads its rows (sph_download_local) and hands them back (sph_upload_local) before every step: the
    ranks recognise the state they hold (one all-reduce of the comparison flags) and end up bit for bit where the same
    domains end up without the round trips."""
    env = {"SPH_TEST_QUIET": "1"}
    ref = run_ranks(tmp_path, f"ddr{world}", world, "host", MODE_VARIABLE_H, steps=4, env_extra=env, n=30_000, domains=1)
    env2 = dict(env, SPH_TEST_ROUNDTRIP="1")
    res = run_ranks(tmp_path, f"ddt{world}", world, "host", MODE_VARIABLE_H, steps=4, env_extra=env2, n=30_000, domains=1)
    for r in range(world):
        assert int(res[r]["resident_hits"][0]) == 3 and int(ref[r]["resident_hits"][0]) == 0, (r, res[r]["resident_hits"])
        for k in ref[r]:
            if k == "resident_hits":
                continue
            a, b = res[r][k], ref[r][k]
            assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), (r, k)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_domains_nccl_two_gpus(reference, tmp_path):
    res = run_ranks(tmp_path, "ddn2", 2, "nccl", MODE_VARIABLE_H, steps=3, domains=1, extra="tree")
    for r in range(2):
        compare(res[r], reference[(MODE_VARIABLE_H, 3)], f"nccl rank {r}", 1e-10, 1e-10)
