"""Development: one factorisation + one solve at size n (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.linalg import CholeskyFactor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
g = torch.randn(n, 256, dtype=torch.float64, device="cuda")
a = g @ g.T / 256 + torch.eye(n, dtype=torch.float64, device="cuda")
f = CholeskyFactor(a, overwrite=True, check=False)
x = f.solve(torch.randn(n, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
print("info", f.info)
