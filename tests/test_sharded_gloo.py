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


# ---------------------------------------------------------------------------------------------------------------
# The sharded public entry point, map_cells_to_space(..., process_group=), on CPU: an oracle-backed stand-in replaces the
# CUDA Mapper, so what is tested is the host contract -- rank-local rows of the reference's M0 stream, per-rank AnnData,
# global per-gene scores, rank-0 gather.
class _FakeShardedMapper:
    """Mapper's sharded surface (process_group, _rows, train, project, release) over OracleShardEngine."""

    class _Cfg:
        device = 0

    def __init__(self, S, G, d=None, device=None, random_state=None, precision=None, process_group=None, lambda_d=0, **kw):
        from tangram_b200.mapping_optimizer import legacy_normal_rows
        self._cfg = self._Cfg()
        self._pg = process_group
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        N, V = S.shape[0], G.shape[0]
        self._rows = shard_rows(N, rank, world)
        r0, r1 = self._rows
        M0 = np.zeros((N, V), dtype=np.float32)
        M0[r0:r1] = legacy_normal_rows(random_state, N, V, r0, r1)
        self.eng = OracleShardEngine(dict(S=S, G=G, d=d), M0, r0, r1, N)
        self.n_cells = r1 - r0

    def train(self, num_epochs, learning_rate=0.1, print_each=None):
        sharded_steps(self.eng, num_epochs, learning_rate, lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._pg))
        out = torch.softmax(self.eng.full.M, dim=1).numpy()
        return out, {k: [float("nan")] * num_epochs for k in ("total_loss", "main_loss", "vg_reg", "kl_reg", "entropy_reg")}

    def project(self, X):
        return (torch.softmax(self.eng.full.M, dim=1).t() @ torch.as_tensor(X)).numpy()

    def release(self):
        pass


def _api_worker(rank, world, port, inp, epochs, out):
    import pandas as pd
    import tangram_b200 as tg
    from tangram_b200 import mapping_utils as mu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    mu.mo.Mapper = _FakeShardedMapper
    N, K = inp["S"].shape
    V = inp["G"].shape[0]
    genes = [f"g{i}" for i in range(K)]
    ad_sc = tg.MiniAnnData(X=inp["S"].copy(), obs=pd.DataFrame({"i": np.arange(N)}, index=[f"c{i}" for i in range(N)]),
                           var=pd.DataFrame(index=genes))
    ad_sp = tg.MiniAnnData(X=inp["G"].copy(), obs=pd.DataFrame(index=[f"v{i}" for i in range(V)]), var=pd.DataFrame(index=genes))
    tg.pp_adatas(ad_sc, ad_sp)
    part = tg.map_cells_to_space(ad_sc, ad_sp, num_epochs=epochs, random_state=9, verbose=False, process_group=dist.group.WORLD)
    full = tg.map_cells_to_space(ad_sc, ad_sp, num_epochs=epochs, random_state=9, verbose=False, process_group=dist.group.WORLD,
                                 gather=True)
    out[rank] = dict(rows=part.uns["shard_rows"], X=np.asarray(part.X), obs=list(part.obs.index),
                     scores=part.uns["train_genes_df"]["train_score"].sort_index().values,
                     full=None if full is None else (np.asarray(full.X), list(full.obs.index)),
                     genes=list(part.uns["train_genes_df"].sort_index().index))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_map_cells_to_space_contract_two_rank_gloo():
    N, V, K, epochs = 37, 11, 9, 5
    inp = synthetic_inputs(N, V, K, seed=3)
    inp["d"] = (inp["G"].sum(axis=1) / inp["G"].sum()).astype(np.float32)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_api_worker, args=(2, port, inp, epochs, out), nprocs=2, join=True)
    # the unsharded oracle from the reference's full draw (random_state=9); the result does not depend on the gene order
    assert out[0]["genes"] == out[1]["genes"]
    order = list(range(K))
    ref = OracleMapper(inp["S"][:, order], inp["G"][:, order], d=inp["d"], lambda_d=1.0, random_state=9)
    ro, _ = ref.train(epochs, print_each=None)
    blocks = sorted((out[r]["rows"], out[r]["X"], out[r]["obs"]) for r in range(2))
    assert blocks[0][0][0] == 0 and blocks[0][0][1] == blocks[1][0][0] and blocks[1][0][1] == N
    got = np.concatenate([b[1] for b in blocks])
    assert np.linalg.norm(got - ro) / np.linalg.norm(ro) < 1e-5
    assert [n for b in blocks for n in b[2]] == [f"c{i}" for i in range(N)]          # each rank: obs of ITS cells
    assert np.allclose(out[0]["scores"], out[1]["scores"], rtol=1e-6)                # per-gene scores are global
    Gp = ro.T @ inp["S"][:, order]
    G = inp["G"][:, order]
    cs = (G * Gp).sum(0) / (np.linalg.norm(G, axis=0) * np.linalg.norm(Gp, axis=0))
    assert np.allclose(np.sort(out[0]["scores"]), np.sort(cs), rtol=1e-4)
    assert out[1]["full"] is None                                                     # gather=True: rank 0 only
    fx, fobs = out[0]["full"]
    assert fx.shape == (N, V) and fobs == [f"c{i}" for i in range(N)]
    assert np.linalg.norm(fx - ro) / np.linalg.norm(ro) < 1e-5
