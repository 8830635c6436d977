"""Dev tool: per-kernel ms of one C3 iteration (tgb200_profile_step), for the product library or a debug build in tools/
named by TGB_DBG_LIB (e.g. built with -DTGB_SKIP_EPI: the tensor-core kernels with their epilogues compiled out)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tangram_b200 import _build, _lib
if os.environ.get("TGB_DBG_LIB"):
    _build.LIB = os.path.join(ROOT, "tools", os.environ["TGB_DBG_LIB"])
    _build.is_current = lambda: True
from tangram_b200.engine import Engine
import bench, numpy as np
N, V, K = 100000, 10000, 2000
inp = bench.gen_inputs("c3", 0, N)
eng = Engine(N, V, K, precision="bf16")
eng.set_expression(inp["S"], inp["G"]); eng.set_density(inp["d"]); eng.init_mapping_normal(1)
eng.run(2)
acc = {}
for _ in range(3):
    step = {}
    for k, v in eng.profile_step():
        step[k] = step.get(k, 0.0) + v          # one launch per cell chunk: add them up
    for k, v in step.items():
        acc.setdefault(k, []).append(v)
print({k: round(float(np.mean(v)), 3) for k, v in acc.items()})
