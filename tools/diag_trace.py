"""Development: phase timeline (SM clocks) of one diag_tile_kernel launch (needs tools/liblsk_trace.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from liestationarykernels_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "liblsk_trace.so")
from liestationarykernels_b200.linalg import CholeskyFactor
torch.manual_seed(0)
n = 128
g = torch.randn(n, 2 * n, dtype=torch.float64, device="cuda")
a = g @ g.T / (2 * n) + torch.eye(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    f = CholeskyFactor(a.clone(), overwrite=True, check=False)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((8, 80), dtype=np.int64)
lib.lsk_debug_diag_trace.restype = ctypes.c_int
lib.lsk_debug_diag_trace.argtypes = [ctypes.c_void_p]
assert lib.lsk_debug_diag_trace(buf.ctypes.data) == 0
b = buf - buf[:, 0].min()
np.set_printoptions(linewidth=250)
# BAR.SYNC does not block the clock read behind it: a post-barrier stamp of an early warp shows its ARRIVAL.  Release time of a
# barrier = latest pre-barrier stamp over the warps.
rel = lambda ev: b[:, ev].max()
print("load done:", rel(1))
prev = rel(1)
for jb in range(8):
    e = 2 + 6 * jb
    a, s2, s3 = rel(e), rel(e + 2), rel(e + 4)
    print("jb %d: S1 | S3b region %5d (warp 0 S1 alone: %5d; S3b warps max %5d)   S2 region %5d   S3a region %5d" % (
        jb, a - prev, b[0, e] - prev, (b[1:, e] - prev).max(), s2 - a, s3 - s2))
    prev = s3
print("factor loop:", prev - rel(1), " last row of the inverse:", rel(55) - rel(50), " total:", rel(55))
