"""Development: tile-kernel throughput for small K (LINEAR / SYRK epilogues, general / triangular schedules)."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import _lib, ops

lib = _lib.load()
res = []
for n, K in [(16384, 128), (16384, 256), (16384, 512), (16384, 4096), (8192, 256)]:
    a = ops.PanelMatrix.from_dense(torch.randn(n, K, dtype=torch.float64, device="cuda"))
    out = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for kind, sym in [(_lib.EPI_LINEAR, 0), (_lib.EPI_LINEAR, 1), (_lib.EPI_SYRK, 0), (_lib.EPI_SYRK, 1)]:
        ep = _lib.LskEpilogue()
        ep.kind, ep.nforms, ep.nvars, ep.nterms = kind, 1, 1, 1
        ep.group_begin[0], ep.group_end[0] = 0, a.kg
        ep.scale = -1.0 if kind == _lib.EPI_SYRK else 1.0
        ep.out_layout = _lib.OUT_ROW_MAJOR
        ep.reserved = 2 if sym else 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = 1e9
        for it in range(4):
            torch.cuda.synchronize()
            ev[0].record()
            _lib.check(lib.lsk_pairwise_poly(_lib.ptr(a.buf), _lib.ptr(a.buf), n, n, a.kg, ctypes.byref(ep), _lib.ptr(out), n,
                                             _lib.stream_ptr()), "pairwise")
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        flops = 2.0 * n * n * K * (0.5 if sym else 1.0)
        rec = {"n": n, "K": K, "kind": "SYRK" if kind == _lib.EPI_SYRK else "LINEAR", "sym": sym, "ms": best, "tflops": flops / best / 1e9}
        res.append(rec)
        print(rec, flush=True)
    del a, out
json.dump(res, open("gpurun_out/syrk_bench.json", "w"), indent=1)
