"""2-GPU NCCL run of the cell-sharded path (needs >= 2 GPUs: run with `gpurun --gpus 2`).  Launched as a
subprocess through torchrun so the single-GPU test session is unaffected."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["TGB_ROOT"])
from oracle.tangram_oracle import OracleMapper, synthetic_inputs
from tangram_b200 import Mapper
from tangram_b200.sharded import shard_rows
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
N, V, K = 3001, 700, 300
inp = synthetic_inputs(N, V, K, seed=5)
M0 = np.random.default_rng(2).standard_normal((N, V)).astype(np.float32)
kw = dict(S=inp["S"], G=inp["G"], d=inp["d"], lambda_d=1.0, lambda_r=1e-3, lambda_g2=0.3)
for prec, tol in (("fp32", 2e-5), ("bf16", 5e-2)):
    m = Mapper(device=f"cuda:{rank}", M0=M0, precision=prec, process_group=dist.group.WORLD, **kw)
    out, hist = m.train(8, print_each=None)
    r0, r1 = shard_rows(N, rank, world)
    assert out.shape == (r1 - r0, V)
    if rank == 0:
        o = OracleMapper(M0=M0, **kw)
        ref, oh = o.train(8, print_each=None)
    else:
        o = OracleMapper(M0=M0, **kw); ref, oh = o.train(8, print_each=None)
    err = np.linalg.norm(out - ref[r0:r1]) / np.linalg.norm(ref[r0:r1])
    dl = max(abs(float(a) - float(b)) for a, b in zip(hist["total_loss"], oh["total_loss"]))
    print(f"rank {rank} {prec}: rel-Frobenius {err:.3e} max loss diff {dl:.3e}", flush=True)
    assert err < tol and dl < (1e-5 if prec == "fp32" else 1e-3)
    assert m._own_comm, "NCCL group: the exchange must run inside tgb200_run on the handle's own communicator"
    m.release()
# the sharded public entry point: every rank passes the same AnnDatas, gets the AnnData of its cells; rank 0 can gather
import pandas as pd
import tangram_b200 as tg
Na, Va, Ka = 1203, 300, 120
ia = synthetic_inputs(Na, Va, Ka, seed=9)
genes = [f"g{i}" for i in range(Ka)]
ad_sc = tg.MiniAnnData(X=ia["S"].copy(), obs=pd.DataFrame(index=[f"c{i}" for i in range(Na)]), var=pd.DataFrame(index=genes))
ad_sp = tg.MiniAnnData(X=ia["G"].copy(), obs=pd.DataFrame(index=[f"v{i}" for i in range(Va)]), var=pd.DataFrame(index=genes))
tg.pp_adatas(ad_sc, ad_sp)
part = tg.map_cells_to_space(ad_sc, ad_sp, device=f"cuda:{rank}", num_epochs=10, random_state=7, verbose=False, precision="fp32",
                             process_group=dist.group.WORLD)
full = tg.map_cells_to_space(ad_sc, ad_sp, device=f"cuda:{rank}", num_epochs=10, random_state=7, verbose=False, precision="fp32",
                             process_group=dist.group.WORLD, gather=True)
d = np.asarray(ad_sp.obs["rna_count_based_density"], dtype=np.float32)
oa = OracleMapper(ia["S"], ia["G"], d=d, lambda_d=1.0, random_state=7)
ra, _ = oa.train(10, print_each=None)
a0, a1 = part.uns["shard_rows"]
assert (a0, a1) == shard_rows(Na, rank, world) and list(part.obs.index) == [f"c{i}" for i in range(a0, a1)]
assert np.linalg.norm(part.X - ra[a0:a1]) / np.linalg.norm(ra[a0:a1]) < 1e-4
assert (full is None) == (rank != 0)
if rank == 0:
    assert full.X.shape == (Na, Va) and np.linalg.norm(full.X - ra) / np.linalg.norm(ra) < 1e-4
    assert len(full.uns["train_genes_df"]) == Ka
dist.barrier()
dist.destroy_process_group()
print("MULTIGPU OK", flush=True)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_sharded_matches_oracle(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TGB_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0 and res.stdout.count("MULTIGPU OK") == 2
