"""Dev tool: where the end-to-end (host buffers in, host buffers out) time of one Mapper call goes at C3."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from tangram_b200 import Mapper, _lib
N, V, K = 100000, 10000, 2000
inp = bench.gen_inputs("c3", 0, N)
M0 = torch.empty((N, V), dtype=torch.float32).pin_memory(); M0.normal_()
Sp = torch.from_numpy(inp["S"]).pin_memory()
torch.cuda.synchronize()
lib = _lib.load()
buf = np.empty((N, V), dtype=np.float32)
t0 = time.perf_counter(); rc = lib.tgb200_host_pin(_lib.ptr(buf), buf.nbytes, 8, 0); t1 = time.perf_counter()
lib.tgb200_host_unpin(_lib.ptr(buf)); t2 = time.perf_counter()
print(f"host_pin of a fresh {buf.nbytes / 1e9:.1f} GB buffer (8 threads): {t1 - t0:.3f} s (rc {rc}), unpin {t2 - t1:.3f} s")
del buf
for steps in (20, 50, 20):
    t0 = time.perf_counter()
    m = Mapper(S=Sp.numpy(), G=inp["G"], d=inp["d"], lambda_d=1.0, M0=M0.numpy(), precision="bf16", device="cuda:0")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out, hist = m.train(steps, print_each=None)
    t2 = time.perf_counter()
    print(f"steps {steps}: ctor+upload {t1 - t0:.3f} s, train+download {t2 - t1:.3f} s, total {t2 - t0:.3f} s -> {steps / (t2 - t0):.1f} it/s; "
          f"rows sum to {float(out[:4].sum(axis=1).mean()):.5f}")
    del m, out
