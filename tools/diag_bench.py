"""Development: time of one diagonal-block launch (lsk_potrf_upper at n = 128 is exactly one diag kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.linalg import CholeskyFactor
torch.manual_seed(0)
for n in (128, 256):
    g = torch.randn(n, 2 * n, dtype=torch.float64, device="cuda")
    a = g @ g.T / (2 * n) + torch.eye(n, dtype=torch.float64, device="cuda")
    works = [a.clone() for _ in range(200)]
    for w in works[:20]:
        CholeskyFactor(w, overwrite=True, check=False)
    works = [a.clone() for _ in range(200)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for w in works:
        f = CholeskyFactor(w, overwrite=True, check=False)
    e1.record()
    torch.cuda.synchronize()
    ref = torch.linalg.cholesky(a)
    print("n=%d  %.2f us per factorisation (host calls included)   err %.2e" % (n, e0.elapsed_time(e1) * 1e3 / 200,
                                                                             ((f.L - ref).abs().max() / ref.abs().max()).item()))
