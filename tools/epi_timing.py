"""Dev tool: per-phase clock64() timing of the backward epilogue.
Build:  nvcc -DTGB_EPI_TIMING ... -o tools/libtiming.so tangram_b200/csrc/tangram_b200.cu
Run on the GPU box:  python tools/epi_timing.py"""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tangram_b200 import _build, _lib
_build.LIB = os.path.join(ROOT, "tools", os.environ.get("TGB_DBG_LIB", "libtiming.so"))
_build.is_current = lambda: True
from tangram_b200.engine import Engine
import bench
N, V, K = 100000, 10000, 2000
inp = bench.gen_inputs("c3", 0, N)
eng = Engine(N, V, K, precision="bf16")
eng.set_expression(inp["S"], inp["G"]); eng.set_density(inp["d"]); eng.init_mapping_normal(1)
eng.run(3)
lib = _lib.load()
out = (ctypes.c_ulonglong * 8)()
lib.tgb200_debug_epi_timing(out, 1)
eng.run(5)
lib.tgb200_debug_epi_timing(out, 0)
v = list(out); n = v[4]
print("warp-tiles", n)
names = ["wait_stage_mbar", "tmem_ld+compute", "fence+pair_barrier", "leader_issue(lane1 view)", "n", "wait_tfull"]
tot = sum(v[i] for i in (0, 1, 2, 3, 5))
for i in (5, 0, 1, 2, 3):
    print(f"{names[i]:18s} {v[i] / n:10.0f} cycles per warp-tile  {100 * v[i] / tot:5.1f}%")
print("total per warp-tile", tot / n)

print("leader issue+wait_read per warp-tile", v[6] / max(v[7], 1))
