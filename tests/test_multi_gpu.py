"""Two ranks on two GPUs of one box (threads in one process, NCCL communicator inside the engine):
the sharded run is bit-identical to the single-GPU run.  Skipped when fewer than 2 GPUs are visible."""
import threading
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_VARIABLE_H, MODE_FIXED_H, ics
from summersph_b200.state import GAS_FIELDS

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", [MODE_VARIABLE_H, MODE_FIXED_H])
def test_two_ranks_bit_identical_to_one(mode, built_engine):
    from summersph_b200.engine import Engine
    p = default_params(mode, bounding_size=95.0)
    b, s = ics.keplerian_disc(60_000, seed=12)
    s.radius[:] = 12.0                      # accretion + bounds removals exercised too
    steps = 3
    with Engine(p, device=0) as e:
        e.upload(b, s)
        dt, t = 0.01, 0.0
        for _ in range(steps):
            dt, t = e.step(dt, t)
        ref_b, ref_s = e.download(); ref = (dt, t, e.sizes(), e.counters())
    world = 2
    engines = [Engine(p, device=r) for r in range(world)]
    uid = engines[0].unique_id()
    out = [None] * world; err = []

    def run(r):
        try:
            e = engines[r]
            e.comm_init(r, world, uid)
            e.upload(b, s)
            dt, t = 0.01, 0.0
            for _ in range(steps):
                dt, t = e.step(dt, t)
            out[r] = (e.download(), (dt, t, e.sizes(), e.counters()))
        except Exception as ex:          # pragma: no cover
            err.append(ex)

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for x in th: x.start()
    for x in th: x.join(timeout=600)
    for e in engines: e.close()
    assert not err, err
    for r in range(world):
        (bb, ss), meta = out[r]
        assert meta[:3] == ref[:3]
        for k in ("density_candidates", "sph_pairs", "grav_accepted"):
            assert meta[3][k] == ref[3][k], k
        for k in GAS_FIELDS:
            assert np.array_equal(getattr(bb, k), getattr(ref_b, k)), (r, k)
        assert np.array_equal(ss.m, ref_s.m) and np.array_equal(ss.x, ref_s.x)
