"""Development: per-warp timeline of CTA 0 of the generic tile kernel on a plain K = 128 Gram (needs tools/liblsk_trace.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from liestationarykernels_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "liblsk_trace.so")
from liestationarykernels_b200 import ops

N, K = 16384, 128
torch.manual_seed(0)
a = ops.PanelMatrix.from_dense(torch.randn(N, K, dtype=torch.float64, device="cuda"))
b = ops.PanelMatrix.from_dense(torch.randn(N, K, dtype=torch.float64, device="cuda"))
out = torch.empty(N, N, dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.gram(a, b, 1.0, out=out)
torch.cuda.synchronize()
buf = np.zeros((16, 10, 16), dtype=np.int64)
issue = np.zeros(128, dtype=np.int64)
lib = _lib.load()
lib.lsk_debug_gtrace.restype = ctypes.c_int
lib.lsk_debug_gtrace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.lsk_debug_gtrace(buf.ctypes.data, issue.ctypes.data) == 0
t0 = buf[:, 0, 0].min()
b_ = buf - t0
iss = issue - t0
C = 4
for tile in (5, 6):
    print("tile", tile, " issue times of its chunks:", [int(v) for v in iss[C * tile: C * tile + C]])
    for w in range(16):
        r = b_[w, tile]
        waits = [int(r[2 + 2 * c] - r[1 + 2 * c]) for c in range(C)]
        comp = [int(r[1 + 2 * (c + 1)] - r[2 + 2 * c]) for c in range(C - 1)] + [int(r[13] - r[2 + 2 * (C - 1)])]
        print("  warp %2d: start %7d | first wait at +%5d | waits %s | chunk compute %s | epilogue %6d" % (
            w, r[0], r[1] - r[0], waits, comp, b_[w, tile + 1, 0] - r[13]))
print("cycles per tile:", (b_[:, 8, 0] - b_[:, 3, 0]) / 5)
