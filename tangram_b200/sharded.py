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
