"""Golden vectors for `MapperConstrained` from the REAL reference (mapping_optimizer.py:411-639), loaded by path."""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.tangram_oracle import synthetic_inputs  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_mo", "/root/reference/tangram/mapping_optimizer.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
torch.set_num_threads(1)

CASES = {
    "constrained_default": dict(N=240, V=90, K=50, epochs=30, seed=21, hyper=dict(target_count=60)),
    "constrained_regs": dict(N=150, V=70, K=30, epochs=25, seed=8,
                             hyper=dict(lambda_d=0.7, lambda_g1=1.0, lambda_g2=0.6, lambda_r=1e-3, lambda_count=0.5,
                                        lambda_f_reg=0.8, target_count=40)),
}
for name, c in CASES.items():
    inp = synthetic_inputs(c["N"], c["V"], c["K"], seed=c["seed"])
    m = ref.MapperConstrained(S=inp["S"], G=inp["G"], d=inp["d"], device="cpu", random_state=c["seed"], **c["hyper"])
    M0, F0 = m.M.detach().numpy().copy(), m.F.detach().numpy().copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out, F_out, hist = m.train(num_epochs=c["epochs"], learning_rate=0.1, print_each=10)
    m1 = ref.MapperConstrained(S=inp["S"], G=inp["G"], d=inp["d"], device="cpu", random_state=c["seed"], **c["hyper"])
    with contextlib.redirect_stdout(io.StringIO()):
        m1.train(num_epochs=1, learning_rate=0.1, print_each=None)

    def fl(s):
        return float(s.split("(")[1].split(",")[0]) if s.startswith("tensor") else float(s)
    save = dict(in_S=inp["S"], in_G=inp["G"], in_d=inp["d"], M0=M0, F0=F0, M1=m1.M.detach().numpy(), F1=m1.F.detach().numpy(),
                output=out, F_out=F_out, printed=np.array(buf.getvalue()), epochs=np.array(c["epochs"]), seed=np.array(c["seed"]),
                total_loss_str0=np.array(hist["total_loss"][0]), main_loss_str0=np.array(hist["main_loss"][0]))
    for k in hist:
        save[k] = np.array([fl(x) for x in hist[k]], dtype=np.float64)
    for k, v in c["hyper"].items():
        save["hp_" + k] = np.array(v, dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **save)
    print(name, os.path.getsize(path) // 1024, "KiB", hist["total_loss"][0], hist["main_loss"][0], hist["count_reg"][0])
