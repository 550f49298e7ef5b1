"""Where a covariance step goes (development aid): embedding launch, pairwise kernel, launch gaps, host overhead.
    python tools/step_breakdown.py [rows] [cols]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from liestationarykernels_b200 import ops
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
torch.manual_seed(0)
space = SO(5, order=100)
kern = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), space).eval()
x, y = space.rand(rows), space.rand(cols)
out = torch.empty((rows, cols), dtype=torch.float64, device="cuda")
plan = space.plan
ep = kern._epilogue(True)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, host * 1e6


ea, eb = ops.embed_pair(plan, x, y)
res = {"rows": rows, "cols": cols}
with torch.no_grad():
    res["embed_pair_us"], res["embed_pair_host_us"] = timed(lambda: ops.embed_pair(plan, x, y))
    res["embed_x_us"], _ = timed(lambda: ops.embed(plan, x, "a"))
    res["embed_y_us"], _ = timed(lambda: ops.embed(plan, y, "b"))
    res["pairwise_us"], res["pairwise_host_us"] = timed(lambda: ops.pairwise_poly(ea, eb, rows, cols, plan.kg_total, ep, out=out))
    res["step_us"], res["step_host_us"] = timed(lambda: kern(x, y, out=out))
    xs, ys = space.rand(64), space.rand(64)
    os_ = torch.empty((64, 64), dtype=torch.float64, device="cuda")
    res["tiny_step_us"], res["tiny_step_host_us"] = timed(lambda: kern(xs, ys, out=os_), reps=200)
print(json.dumps(res))
