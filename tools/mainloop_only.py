"""Dev tool: time the three tensor-core kernels with their epilogues compiled out (-DTGB_SKIP_EPI)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tangram_b200 import _build, _lib
_build.LIB = os.path.join(ROOT, "tools", os.environ.get("TGB_DBG_LIB", "libtiming.so"))
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
    for k, v in eng.profile_step():
        acc.setdefault(k, []).append(v)
print({k: round(float(np.mean(v)), 3) for k, v in acc.items()})
