"""Development: per-warp timeline of CTA 0 of the C2 pairwise kernel (needs tools/liblsk_trace.so)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from liestationarykernels_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "liblsk_trace.so")
from liestationarykernels_b200 import ops
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure

N = 16384
sp = SO(5, order=100)
k = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), sp); k.eval()
torch.manual_seed(0)
x, y = sp.rand(N), sp.rand(N)
plan = sp.plan
ea, eb = ops.embed_pair(plan, x, y)
out = torch.empty(N, N, dtype=torch.float64, device="cuda")
ep = k._epilogue(True)
for _ in range(3):
    ops.pairwise_poly(ea, eb, N, N, plan.kg_total, ep, out=out)
torch.cuda.synchronize()
buf = np.zeros((16, 12, 8), dtype=np.int64)
lib = _lib.load()
lib.lsk_debug_trace.restype = ctypes.c_int
lib.lsk_debug_trace.argtypes = [ctypes.c_void_p]
assert lib.lsk_debug_trace(buf.ctypes.data) == 0
t0 = buf[:, 0, 0].min()
b = buf - t0
names = ["tile", "data", "dmma_end", "E0", "E0end", "E1", "E1end"]
for tile in range(6, 7):
    print("tile", tile)
    for w in range(16):
        r = b[w, tile]
        print("  warp %2d (smsp %d): start %7d | wait %5d | dmma %5d | var0 %5d | horner0 %5d | st+var1 %5d | horner1 %5d | tail->next %5d" % (
            w, w % 4, r[0], r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], b[w, tile + 1, 0] - r[6]))
print("issue time of chunk 0 of tile k (warp 0's clock) minus start of tile k:", [int(b[0, k, 7] - b[0, k, 0]) for k in range(1, 10)])
per = b[:, 8, 0] - b[:, 3, 0]
print("cycles per tile (tiles 3..8):", per / 5)
