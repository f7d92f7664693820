"""Host-side multi-rank logic on CPU: world_size-2 gloo process group (no GPU): the slice partition is a
disjoint cover and every rank derives the same bounds; the communicator bootstrap broadcasts rank 0's id."""
import os
import sys
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEngine:
    def __init__(self, rank):
        self.rank, self.got = rank, None
    def unique_id(self):
        return bytes([7]) * 128
    def comm_init(self, rank, world, uid):
        self.got = (rank, world, uid)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from summersph_b200.parallel import init_comm, torch_broadcast_bytes, slice_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = FakeEngine(rank)
    init_comm(e, rank, world, torch_broadcast_bytes)
    # every rank computes all slices; all-gather them and check agreement
    mine = torch.tensor([b for r in range(world) for b in slice_bounds(703381, r, world)])
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    same = all(bool((o == mine).all()) for o in out)
    # a dt-style min all-reduce and a sink-style sum all-reduce behave as the engine assumes
    v = torch.tensor([0.01 * (rank + 1)], dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.MIN)
    s = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(s, op=dist.ReduceOp.SUM)
    q.put((rank, e.got, same, float(v), float(s)))
    dist.destroy_process_group()


def test_world_size_2_bootstrap_and_slices():
    world, port = 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, same, vmin, ssum in res:
        assert got == (rank, world, bytes([7]) * 128)
        assert same and vmin == 0.01 and ssum == 3.0


def test_slice_bounds_cover():
    from summersph_b200.parallel import slice_bounds
    for ng in (1, 7, 32, 703381):
        for world in (1, 2, 4, 8):
            b = [slice_bounds(ng, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == ng
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(g1 >= g0 for g0, g1 in b)
            assert all(g0 % 64 == 0 for g0, _ in b)          # cuts only where the gravity runs restart (GRAV_SEG)
