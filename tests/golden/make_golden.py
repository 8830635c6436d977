"""
Generate golden vectors from the REAL reference implementation.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

It loads /root/reference/tangram/mapping_optimizer.py *by path* (the package import
fails here: scanpy is absent), runs the reference `Mapper` on small seeded inputs on
the CPU and stores inputs + outputs in tests/golden/*.npz.  The fixtures travel to the
GPU box; the reference does not.
"""
import importlib.util
import io
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.tangram_oracle import synthetic_inputs, grid_graph, spatial_weights_from_graph  # noqa: E402

REF = "/root/reference/tangram/mapping_optimizer.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_mapping_optimizer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = {
    # name: (N, V, K, T, epochs, seed, clusters, hyper)
    "cells_default": dict(N=300, V=120, K=60, T=0, epochs=30, seed=42, clusters=False,
                          hyper=dict(lambda_g1=1.0, lambda_d=1.0)),
    "cells_nodensity": dict(N=64, V=48, K=24, T=0, epochs=12, seed=7, clusters=False, no_d=True,
                            hyper=dict(lambda_g1=1.0)),
    "cells_regs": dict(N=160, V=90, K=40, T=0, epochs=20, seed=3, clusters=False,
                       hyper=dict(lambda_g1=1.0, lambda_d=0.7, lambda_g2=0.5, lambda_r=1e-3,
                                  lambda_l1=1e-6, lambda_l2=1e-5)),
    "clusters": dict(N=16, V=150, K=50, T=0, epochs=40, seed=11, clusters=True,
                     hyper=dict(lambda_g1=1.0, lambda_d=1.0)),
    "cells_spatial": dict(N=200, V=100, K=40, T=5, epochs=20, seed=5, clusters=False,
                          hyper=dict(lambda_g1=1.0, lambda_d=1.0, lambda_g2=0.3, lambda_r=2.95e-9,
                                     lambda_l2=1e-18, lambda_neighborhood_g1=0.96,
                                     lambda_ct_islands=0.17, lambda_getis_ord=0.71)),
}


def build_case(name):
    c = CASES[name]
    inp = synthetic_inputs(c["N"], c["V"], c["K"], seed=c["seed"], n_types=c["T"], clusters=c["clusters"])
    kw = dict(S=inp["S"], G=inp["G"], d=None if c.get("no_d") else inp["d"])
    if c["clusters"]:
        kw["d_source"] = inp["d_source"]
    hyper = dict(c["hyper"])
    extra = {}
    if hyper.get("lambda_neighborhood_g1", 0) > 0 or hyper.get("lambda_ct_islands", 0) > 0 \
            or hyper.get("lambda_getis_ord", 0) > 0:
        conn, dist = grid_graph(c["V"])
        # mapping_utils.py:319-329
        extra["voxel_weights"] = spatial_weights_from_graph(conn, dist, True, True).toarray()
        extra["neighborhood_filter"] = spatial_weights_from_graph(conn, dist, False, False).toarray()
        extra["spatial_weights"] = spatial_weights_from_graph(conn, dist, False, True).toarray()
        extra["ct_encode"] = inp["ct_encode"]
    kw.update(hyper)
    kw.update(extra)
    return c, kw


def main():
    ref = load_reference()
    torch.set_num_threads(1)  # fixed summation order for the stored vectors
    for name in CASES:
        c, kw = build_case(name)
        seed = c["seed"]
        mapper = ref.Mapper(device="cpu", random_state=seed, **kw)
        M0 = mapper.M.detach().numpy().copy()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out, hist = mapper.train(num_epochs=c["epochs"], learning_rate=0.1, print_each=10)
        # 1-step state for tight checks
        mapper1 = ref.Mapper(device="cpu", random_state=seed, **kw)
        with contextlib.redirect_stdout(io.StringIO()):
            mapper1.train(num_epochs=1, learning_rate=0.1, print_each=None)
        M1 = mapper1.M.detach().numpy().copy()
        save = dict(
            M0=M0, M1=M1, M_final=mapper.M.detach().numpy(), output=out,
            total_loss=np.array([float(x) for x in hist["total_loss"]], dtype=np.float64),
            main_loss=np.array(hist["main_loss"], dtype=np.float64),
            vg_reg=np.array(hist["vg_reg"], dtype=np.float64),
            kl_reg=np.array(hist["kl_reg"], dtype=np.float64),
            entropy_reg=np.array(hist["entropy_reg"], dtype=np.float64),
            printed=np.array(buf.getvalue()),
            epochs=np.array(c["epochs"]), seed=np.array(seed),
        )
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                save["in_" + k] = v
            elif v is not None:
                save["hp_" + k] = np.array(v, dtype=np.float64)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **save)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB",
              "final total_loss", save["total_loss"][-1])


if __name__ == "__main__":
    main()
