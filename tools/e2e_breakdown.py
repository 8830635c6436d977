"""Dev tool: where the end-to-end (host buffers in, host buffers out) time of one Mapper call goes at C3."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from tangram_b200 import Mapper
N, V, K = 100000, 10000, 2000
inp = bench.gen_inputs("c3", 0, N)
M0 = torch.empty((N, V), dtype=torch.float32).pin_memory(); M0.normal_()
Sp = torch.from_numpy(inp["S"]).pin_memory()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    m = Mapper(S=Sp.numpy(), G=inp["G"], d=inp["d"], lambda_d=1.0, M0=M0.numpy(), precision="bf16", device="cuda:0")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    import tangram_b200.mapping_optimizer as mo
    tp0 = time.perf_counter(); buf = mo._pinned_empty((N, V)); tp1 = time.perf_counter(); del buf
    out, hist = m.train(20, print_each=None)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: ctor {t1 - t0:.3f} s, pinned alloc of output {tp1 - tp0:.3f} s, train(20)+download {t2 - t1 - (tp1 - tp0):.3f} s")
    del m, out
