"""fp64 Gram (A B^T) throughput: lsk pairwise kernel (DMMA, EPI_LINEAR) vs cuBLAS DGEMM through torch.matmul."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import ops


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = []
for (n, m, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (16384, 16384, 4096), (8192, 8192, 20480)]:
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    b = torch.randn(m, k, dtype=torch.float64, device="cuda")
    pa, pb = ops.PanelMatrix.from_dense(a), ops.PanelMatrix.from_dense(b)
    out = torch.empty(n, m, dtype=torch.float64, device="cuda")
    t_lsk = timeit(lambda: ops.gram(pa, pb, 1.0, out=out))
    ref = None
    t_cublas = timeit(lambda: torch.matmul(a, b.T, out=out))
    err = (ops.gram(pa, pb, 1.0) - a @ b.T).abs().max().item()
    fl = 2.0 * n * m * k
    res.append(dict(n=n, m=m, k=k, lsk_ms=t_lsk, cublas_ms=t_cublas, lsk_tflops=fl / t_lsk / 1e9, cublas_tflops=fl / t_cublas / 1e9, max_abs_err=err))
    print(res[-1], flush=True)
json.dump(res, open("gpurun_out/gram_bench.json", "w"), indent=1)
