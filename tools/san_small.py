"""Dev tool: a few iterations of every precision mode at shapes that hit the CTA-pair backward kernels (ragged in both
dimensions, second CTA of the last pair fully out of range) and the single-CTA kernels -- for compute-sanitizer memcheck.
With TGB200_CHUNKS=2 in the environment the 2100-cell case also runs the chunked three-stream pipeline (store-only backward
contraction per chunk, streaming Adam on the update stream, next forward's chunks on the third stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.tangram_oracle import synthetic_inputs
from tangram_b200 import Mapper
for (N, V, K) in ((2100, 300, 70), (130, 65, 3)):
    inp = synthetic_inputs(N, V, K, seed=3)
    for prec in ("bf16", "bf16x3", "fp32"):
        m = Mapper(S=inp["S"], G=inp["G"], d=inp["d"], lambda_g1=1.0, lambda_d=1.0, lambda_g2=0.5, lambda_r=1e-3, precision=prec, device="cuda:0")
        out, hist = m.train(4, print_each=None)
        assert np.all(np.isfinite(out)), (N, V, K, prec)
        print(N, V, K, prec, float(hist["total_loss"][-1]), flush=True)
print("done")
