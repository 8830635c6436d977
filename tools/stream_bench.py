"""Dev tool: practical HBM ceiling of the backward epilogue's access mix (3 fp32 streams in, 3 fp32 + 1 bf16 out)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tangram_b200 import _build, _lib
_build.LIB = os.path.join(ROOT, "tools", "libtiming.so")
_build.is_current = lambda: True
from tangram_b200.engine import Engine
import numpy as np
N, V, K = 100000, 10000, 64
eng = Engine(N, V, K, precision="bf16")
S = np.ones((N, K), np.float32); G = np.ones((V, K), np.float32)
eng.set_expression(S, G); eng.set_density((np.ones(V) / V).astype(np.float32)); eng.init_mapping_normal(1)
lib = _lib.load()
ms = ctypes.c_float()
for bps in (2, 4, 8, 16):
    lib.tgb200_debug_stream_bench(eng._h, bps, ctypes.byref(ms))
    gb = N * 10048 * 26 / 1e9
    print(f"blocks/SM {bps:2d}: {ms.value:.3f} ms  -> {gb / ms.value:.1f} GB/s... ({gb:.1f} GB)")
