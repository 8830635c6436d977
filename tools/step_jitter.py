"""Dev tool: per-step device time over a long run vs the sum of per-kernel times from profile_step (C3, bf16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from tangram_b200.engine import Engine
N, V, K = 100000, 10000, 2000
inp = bench.gen_inputs("c3", 0, N)
eng = Engine(N, V, K, precision="bf16")
eng.set_expression(inp["S"], inp["G"]); eng.set_density(inp["d"]); eng.init_mapping_normal(1)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3): eng.run(1, 0.1, stream)
torch.cuda.synchronize()
n = 60
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    eng.run(1, 0.1, stream); ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
print("per-step ms: first 10", np.round(ms[:10], 2), " mean[10:]", ms[10:].mean().round(3), " min", ms.min().round(3), " max", ms.max().round(3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.run(40, 0.1, stream); e1.record(); torch.cuda.synchronize()
print("run(40) in one call: ms/step", round(e0.elapsed_time(e1) / 40, 3))
tot = []
for _ in range(5):
    tot.append(sum(v for _, v in eng.profile_step(0.1, stream)))
print("profile_step sums:", np.round(tot, 3))
