"""Development: per-warp timeline of CTA 0 of the generic pair kernel on the SU(4) config (needs tools/liblsk_trace.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from liestationarykernels_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "liblsk_trace.so")
from liestationarykernels_b200 import ops
from liestationarykernels_b200.spaces import SU
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure

N = 16384
sp = SU(4, order=20)
k = EigenbasisSumKernel(SqExpSpectralMeasure(15, 1.0), sp); k.eval()
torch.manual_seed(0)
x, y = sp.rand(N), sp.rand(N)
plan = sp.plan
ea, eb = ops.embed_pair(plan, x, y)
out = torch.empty(N, N, dtype=torch.float64, device="cuda")
ep = k._epilogue(True)
for _ in range(2):
    ops.pairwise_poly(ea, eb, N, N, plan.kg_total, ep, out=out)
torch.cuda.synchronize()
buf = np.zeros((16, 10, 16), dtype=np.int64)
issue = np.zeros(128, dtype=np.int64)
lib = _lib.load()
lib.lsk_debug_gtrace.restype = ctypes.c_int
lib.lsk_debug_gtrace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.lsk_debug_gtrace(buf.ctypes.data, issue.ctypes.data) == 0
t0 = buf[:, 0, 0].min()
b = buf - t0
iss = issue - t0
for tile in (6,):
    print("tile", tile, " issue times of its 5 chunks:", [int(v) for v in iss[5 * tile: 5 * tile + 5]])
    for w in range(16):
        r = b[w, tile]
        waits = [int(r[2 + 2 * c] - r[1 + 2 * c]) for c in range(5)]
        comp = [int(r[1 + 2 * (c + 1)] - r[2 + 2 * c]) for c in range(4)] + [int(r[13] - r[10])]
        print("  warp %2d: start %7d | wait-begin of chunks %s | waits %s | chunk compute %s | epilogue %6d" % (
            w, r[0], [int(r[1 + 2 * c]) for c in range(5)], waits, comp, b[w, tile + 1, 0] - r[13]))
print("cycles per tile:", (b[:, 8, 0] - b[:, 3, 0]) / 5)
