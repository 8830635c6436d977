"""world_size-2 gloo test (CPU) of the cell-sharded loop in tangram_b200/sharded.py: an oracle-backed
engine stands in for the CUDA handle, so what is tested is the host protocol -- row partition, the
contents of the exchange buffer, one all-reduce per step -- against the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.tangram_oracle import OracleMapper, synthetic_inputs
from tangram_b200.sharded import shard_rows, sharded_steps


def test_shard_rows_partition():
    for n, w in ((10, 3), (100000, 8), (7, 8), (256, 1)):
        blocks = [shard_rows(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


class OracleShardEngine:
    """step_begin / exchange_tensor / step_end over the oracle's math for rows [r0, r1)."""

    def __init__(self, inp, M0, r0, r1, n_global, lam_r=0.0):
        self.S = torch.as_tensor(inp["S"][r0:r1])
        self.G = torch.as_tensor(inp["G"])
        self.d = torch.as_tensor(inp["d"])
        self.full = OracleMapper(inp["S"][r0:r1], inp["G"], d=inp["d"], lambda_d=1.0, lambda_r=lam_r, M0=M0[r0:r1])
        self.n_global = n_global
        V, K = self.G.shape
        self.buf = torch.zeros(V * (K + 1) + 4)
        self.lam_r = lam_r

    def exchange_tensor(self):
        return self.buf

    def step_begin(self):
        P = torch.softmax(self.full.M, dim=1)
        V, K = self.G.shape
        Yx = torch.cat([P.t() @ self.S, P.sum(dim=0, keepdim=True).t()], dim=1)   # [Y | column sums]
        self.buf[: V * (K + 1)] = Yx.reshape(-1)
        self.buf[V * (K + 1)] = (P * torch.log(P)).sum()                            # entropy partial
        self.P = P

    def step_end(self, lr):
        V, K = self.G.shape
        Yx = self.buf[: V * (K + 1)].reshape(V, K + 1)
        Y, cs = Yx[:, :K], Yx[:, K]
        G, d = self.G, self.d
        ny = torch.linalg.vector_norm(Y, dim=0).clamp_min(1e-8)
        ng = torch.linalg.vector_norm(G, dim=0).clamp_min(1e-8)
        c = (Y * G).sum(0) / (ny * ng)
        dY = -(G / (ny * ng) - Y * (c / ny ** 2)) / K
        gd = -d / cs                                            # d KL / d colsum (n_global cancels)
        dP = self.S @ dY.t() + gd[None, :]
        if self.lam_r:
            dP = dP - self.lam_r * (torch.log(self.P) + 1.0)
        P = self.P
        dM = P * (dP - (P * dP).sum(1, keepdim=True))
        self.full.adam_step(dM, lr)


def _worker(rank, world, port, inp, M0, steps, lam_r, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    r0, r1 = shard_rows(inp["S"].shape[0], rank, world)
    eng = OracleShardEngine(inp, M0, r0, r1, inp["S"].shape[0], lam_r)
    sharded_steps(eng, steps, 0.1, lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
    out[rank] = (r0, r1, eng.full.M.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lam_r", [0.0, 1e-3])
def test_two_rank_gloo_matches_unsharded_oracle(lam_r):
    N, V, K, steps = 203, 40, 30, 6                      # odd N: uneven shards
    inp = synthetic_inputs(N, V, K, seed=13)
    M0 = np.random.default_rng(1).standard_normal((N, V)).astype(np.float32)
    ref = OracleMapper(inp["S"], inp["G"], d=inp["d"], lambda_d=1.0, lambda_r=lam_r, M0=M0)
    for _ in range(steps):
        _, g = ref.loss_and_grad()
        ref.adam_step(g, 0.1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, inp, M0, steps, lam_r, out), nprocs=2, join=True)
    got = np.zeros_like(M0)
    for r in range(2):
        r0, r1, Mr = out[r]
        got[r0:r1] = Mr
    assert np.linalg.norm(got - ref.M.numpy()) / np.linalg.norm(ref.M.numpy()) < 1e-5
