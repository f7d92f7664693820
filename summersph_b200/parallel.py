"""Multi-GPU plumbing: one process per GPU, the communicator (NCCL, or host shared memory for tests) owned by the engine.

Two forms, chosen by `params.decomposition`:
  1 - Morton-ordered domains: every rank owns a contiguous range of the global descent-key order plus a halo; top tree by
      all-gather, locally essential tree pulled from the peers, migration at every tree build; results equal the
      single-rank run to rounding (order, keys, leaf cells, neighbour sets and counters bit for bit).  `bench.py --gpus N`.
  0 - replicated state: every rank holds all particles and builds the whole tree, the target work is sharded by
      contiguous Morton slices of walk groups (`slice_bounds`), results travel by peer-memory pushes; bit-identical to
      the single-rank run.
"""


def slice_bounds(n_groups, rank, world):
    """Group slice [g0, g1) of `rank`: the engine's own `sph_slice_bounds` (the function compute_slices() in
    sph_engine.cu uses; cuts fall on multiples of 64 groups so the gravity runs do not depend on the rank count)."""
    import ctypes as C
    from .engine import load_library
    first = (C.c_int32 * (world + 1))()
    rc = load_library().sph_slice_bounds(int(n_groups), int(world), first)
    if rc != 0:
        raise ValueError("sph_slice_bounds: bad arguments")
    return int(first[rank]), int(first[rank + 1])


def init_comm(engine, rank, world, broadcast_bytes):
    """Create the engine's NCCL communicator. `broadcast_bytes(b: bytes|None) -> bytes` must return rank 0's
    payload on every rank (torch.distributed, MPI, a pipe ...)."""
    if world <= 1:
        return
    uid = engine.unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid)
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise ValueError("unique id broadcast failed")
    engine.comm_init(rank, world, bytes(uid))


def torch_broadcast_bytes(payload, src=0):
    """broadcast_bytes implementation on top of an initialised torch.distributed process group."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]
