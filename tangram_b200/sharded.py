"""Cell-sharded operation: one process per GPU, each rank owns a contiguous block of
cell rows of M / S / Adam state; the only cross-rank coupling per iteration is the sum over cells in
Y_ext = sum_r P_r^T S_ext,r (predicted expression | density columns | cell-type columns) plus three
scalars (entropy, L1, L2 partial sums).  So: ONE sum-all-reduce of the exchange buffer per step.

The loop is written against a tiny engine protocol so the same code drives the CUDA handle (NCCL)
and, in the CPU tests, an oracle-backed stand-in (gloo):
    engine.step_begin()           -> local partial sums are in engine.exchange_tensor()
    all_reduce(tensor)            -> in-place sum over ranks
    engine.step_end(lr)           -> loss, backward, Adam on the local rows
"""


def shard_rows(n_cells, rank, world):
    """Contiguous cell-row block of `rank`: balanced to within one row."""
    base, rem = divmod(n_cells, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def sharded_steps(engine, n_steps, lr, all_reduce):
    """Run `n_steps` iterations of the cell-sharded loop."""
    buf = engine.exchange_tensor()
    for _ in range(n_steps):
        engine.step_begin()
        all_reduce(buf)
        engine.step_end(lr)


# ---- NCCL communicators for the in-library exchange, one per (process group, device), kept for the life of the process ----
_COMMS = {}


def nccl_comm_for_group(pg, device):
    """(ncclComm_t, rank, world) for a torch.distributed NCCL group: created on first use (rank 0 makes the 128-byte
    unique id with tgb200_comm_unique_id, one dist.broadcast carries it), then reused by every Mapper / Engine of this
    process -- ncclCommInitRank takes a second or more at 8 ranks.  Returns None for non-NCCL groups (gloo in the CPU tests),
    which keep the host-driven exchange of `sharded_steps`.  The communicators are deliberately not destroyed at interpreter
    exit (handles that borrowed one may be finalised later; the process is going away anyway)."""
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    from . import _lib
    if dist.get_backend(pg) != "nccl":
        return None
    key = (id(pg), int(device))
    if key not in _COMMS:
        lib = _lib.load()
        rank, world = dist.get_rank(pg), dist.get_world_size(pg)
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(lib.tgb200_comm_unique_id(_lib.ptr(uid), uid.nbytes))
        t = torch.from_numpy(uid).to(f"cuda:{int(device)}")
        dist.broadcast(t, src=dist.get_global_rank(pg, 0), group=pg)
        uid = np.ascontiguousarray(t.cpu().numpy())
        comm = ctypes.c_void_p()
        _lib.check(lib.tgb200_comm_create(_lib.ptr(uid), rank, world, int(device), ctypes.byref(comm)))
        _COMMS[key] = (comm, rank, world, pg)          # pg kept alive: id(pg) stays unique
    comm, rank, world, _ = _COMMS[key]
    return comm, rank, world
