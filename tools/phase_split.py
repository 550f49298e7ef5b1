"""Where does the C2 pairwise kernel spend its time?  (a) the full kernel, (b) its DMMA phase alone (LINEAR epilogue on
the same K=128 operands), (c) its epilogue alone (same polynomial, operands cut to one k-group per form)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import _lib, ops
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


N = 16384
sp = SO(5, order=100)
k = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), sp); k.eval()
torch.manual_seed(0)
x, y = sp.rand(N), sp.rand(N)
plan = sp.plan
ea, eb = ops.embed(plan, x, "a"), ops.embed(plan, y, "b")
out = torch.empty(N, N, dtype=torch.float64, device="cuda")
ep = k._epilogue(True)
res = {}
res["full_ms"] = timeit(lambda: ops.pairwise_poly(ea, eb, N, N, plan.kg_total, ep, out=out))
lin = _lib.LskEpilogue()
lin.kind, lin.nforms, lin.nvars, lin.nterms = _lib.EPI_LINEAR, 1, 1, 1
lin.group_begin[0], lin.group_end[0] = 0, plan.kg_total
lin.scale = 1.0
res["dmma_only_ms"] = timeit(lambda: ops.pairwise_poly(ea, eb, N, N, plan.kg_total, lin, out=out))
# epilogue only: 2 k-groups total (one per form)
small = torch.zeros(ops._lib.load().lsk_embed_buffer_doubles(N, 2), dtype=torch.float64, device="cuda")
small.normal_()
ep2 = _lib.LskEpilogue.from_buffer_copy(bytes(ep))
ep2.group_begin[0], ep2.group_end[0], ep2.group_begin[1], ep2.group_end[1] = 0, 1, 1, 2
res["epilogue_only_ms"] = timeit(lambda: ops.pairwise_poly(small, small, N, N, 2, ep2, out=out))
lin2 = _lib.LskEpilogue.from_buffer_copy(bytes(lin))
lin2.group_end[0] = 2
res["store_only_ms"] = timeit(lambda: ops.pairwise_poly(small, small, N, N, 2, lin2, out=out))
print(json.dumps(res))
res2 = {}
for rows in (1, 3, 6, 11):
    e = _lib.LskEpilogue.from_buffer_copy(bytes(ep2))
    e.nterms = rows
    e.stair_shape = ep2.stair_shape & ((1 << (4 * rows)) - 1)
    steps = sum(e.row_len[i] for i in range(rows)) // 2
    res2["epilogue_%drows_%dsteps_ms" % (rows, steps)] = timeit(lambda: ops.pairwise_poly(small, small, N, N, 2, e, out=out))
print(json.dumps(res2))
