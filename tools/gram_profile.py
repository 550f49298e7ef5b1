"""Development: one LINEAR Gram (A B^T, n = m = 8192, K = 16384 by default) for ncu captures of the tile kernel at large K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
torch.manual_seed(0)
a = torch.randn(n, k, dtype=torch.float64, device="cuda")
b = torch.randn(n, k, dtype=torch.float64, device="cuda")
pa, pb = ops.PanelMatrix.from_dense(a), ops.PanelMatrix.from_dense(b)
del a, b
out = torch.empty(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.gram(pa, pb, 1.0, out=out)
torch.cuda.synchronize()
print("ok", out.shape)
