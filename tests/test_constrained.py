"""`MapperConstrained` (mapping_optimizer.py:411-639): oracle pinned to golden vectors from the live
reference (CPU), CUDA path against the same vectors in both parity-grade modes (GPU)."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle.tangram_oracle import OracleMapperConstrained
from tests.helpers import GOLDEN_DIR, assert_same_print, max_rel, rel_fro

CASES = ["constrained_default", "constrained_regs"]
KEYS = ["total_loss", "main_loss", "vg_reg", "kl_reg", "entropy_reg", "count_reg", "lambda_f_reg"]


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    hp = {k[3:]: float(z[k]) for k in z.files if k.startswith("hp_")}
    return z, hp


def _floats(hist, key):
    return np.array([float(s.split("(")[1].split(",")[0]) if s.startswith("tensor") else float(s) for s in hist[key]])


def _check_run(z, out, F_out, hist, printed):
    assert max_rel(_floats(hist, "total_loss"), z["total_loss"]) < 1e-4
    for k in KEYS[1:]:
        a, b = _floats(hist, k), z[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        if not np.isnan(b).all():
            # kl_reg is a small difference of O(1) sums: compare with an absolute floor of fp32 cancellation
            assert np.all(np.abs(a - b)[~np.isnan(b)] <= 2e-4 * np.abs(b[~np.isnan(b)]) + 1e-6), k
    assert rel_fro(out, z["output"]) < 1e-4
    assert rel_fro(F_out, z["F_out"]) < 1e-4
    assert all(isinstance(x, str) for k in KEYS for x in hist[k])                 # :630
    assert hist["total_loss"][0].startswith("tensor(") and hist["total_loss"][0].endswith("grad_fn=<AddBackward0>)")
    ours, ref = printed.splitlines(), str(z["printed"]).splitlines()
    assert len(ours) == len(ref)
    assert_same_print(ours[0], ref[0])


@pytest.mark.parametrize("name", CASES)
def test_constrained_oracle_matches_reference_golden(name):
    z, hp = _load(name)
    o = OracleMapperConstrained(z["in_S"], z["in_G"], z["in_d"], random_state=int(z["seed"]), **hp)
    assert np.array_equal(o.M.numpy(), z["M0"]) and np.array_equal(o.F.numpy(), z["F0"])    # draw order (:472-493)
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        out, F_out, hist = o.train(int(z["epochs"]), print_each=10)
    _check_run(z, out, F_out, hist, buf.getvalue())
    assert hist["total_loss"][0] == str(z["total_loss_str0"]) and hist["main_loss"][0][:8] == str(z["main_loss_str0"])[:8]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", CASES)
def test_constrained_cuda_matches_reference_golden(name, precision):
    from tangram_b200 import MapperConstrained
    z, hp = _load(name)
    m1 = MapperConstrained(z["in_S"], z["in_G"], z["in_d"], device="cuda:0", random_state=int(z["seed"]), precision=precision, **hp)
    m1.train(1, print_each=None)
    M1, F1, step = m1.state()
    assert step == 1 and rel_fro(M1, z["M1"]) < 1e-5 and rel_fro(F1, z["F1"]) < 1e-5
    m = MapperConstrained(z["in_S"], z["in_G"], z["in_d"], device="cuda:0", random_state=int(z["seed"]), precision=precision, **hp)
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        out, F_out, hist = m.train(int(z["epochs"]), print_each=10)
    _check_run(z, out, F_out, hist, buf.getvalue())


@pytest.mark.gpu
def test_constrained_bf16_tracks_oracle():
    from oracle.tangram_oracle import synthetic_inputs
    from tangram_b200 import MapperConstrained
    inp = synthetic_inputs(1500, 400, 200, seed=4)
    o = OracleMapperConstrained(inp["S"], inp["G"], inp["d"], target_count=300, random_state=6)
    M0, F0 = o.M.numpy().copy(), o.F.numpy().copy()
    oo, oF, _ = o.train(20, print_each=None)
    m = MapperConstrained(inp["S"], inp["G"], inp["d"], target_count=300, device="cuda:0", precision="bf16", M0=M0, F0=F0)
    out, F_out, hist = m.train(20, print_each=None)
    assert max_rel(_floats(hist, "total_loss"), o.float_history["total_loss"]) < 2e-3
    assert rel_fro(F_out, oF) < 2e-2 and rel_fro(out, oo) < 5e-2


@pytest.mark.gpu
def test_map_cells_to_space_constrained_mode():
    import pandas as pd
    import tangram_b200 as tg
    from oracle.tangram_oracle import synthetic_inputs
    N, V, K = 200, 60, 40
    inp = synthetic_inputs(N, V, K, seed=12)
    genes = [f"Gene{i}" for i in range(K)]
    ad_sc = tg.MiniAnnData(X=inp["S"].copy(), obs=pd.DataFrame(index=[f"c{i}" for i in range(N)]), var=pd.DataFrame(index=genes))
    ad_sp = tg.MiniAnnData(X=inp["G"].copy(), obs=pd.DataFrame(index=[f"v{i}" for i in range(V)]), var=pd.DataFrame(index=genes))
    tg.pp_adatas(ad_sc, ad_sp)
    ad_map = tg.map_cells_to_space(ad_sc, ad_sp, mode="constrained", target_count=50, lambda_f_reg=1, lambda_count=1,
                                   device="cuda:0", num_epochs=15, random_state=3, verbose=False)
    assert ad_map.X.shape == (N, V) and "F_out" in ad_map.obs.keys()
    tr = ad_sc.uns["training_genes"]
    S = np.asarray(ad_sc[:, tr].X, dtype=np.float32)
    G = np.asarray(ad_sp[:, tr].X, dtype=np.float32)
    o = OracleMapperConstrained(S, G, np.asarray(ad_sp.obs["rna_count_based_density"], dtype=np.float32), lambda_d=1, lambda_g1=1,
                                lambda_g2=0, lambda_r=0, lambda_count=1, lambda_f_reg=1, target_count=50, random_state=3)
    oo, oF, _ = o.train(15, print_each=None)
    assert rel_fro(ad_map.X, oo) < 1e-4 and rel_fro(np.asarray(ad_map.obs["F_out"]), oF) < 1e-4
