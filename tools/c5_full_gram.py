"""Measurement: the full C5 Gram on ONE GPU (H^3, N = 65536, m = 8192 spherical functions: 65536^2 x 16384, 34.4 GB result)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import ops
from liestationarykernels_b200.spaces import HyperbolicSpace
from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure

torch.manual_seed(0)
N, m = 65536, 8192
space = HyperbolicSpace(3, order=m)
kern = RandomFourierFeatureKernel(MaternSpectralMeasure(3, 1.0, 2.5), space).eval()
x, y = space.rand(N), space.rand(N)
out = torch.empty(N, N, dtype=torch.float64, device="cuda")
res = {}
with torch.no_grad():
    for name, args in (("general_x_y", (x, y)), ("symmetric_x_x", (x,))):
        lazy = kern(*args)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = 1e9
        for it in range(2):
            torch.cuda.synchronize()
            ev[0].record()
            lazy.evaluate(out=out)
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        flops = 2.0 * N * N * 2 * m
        res[name] = {"ms": best, "tflops_full_problem": flops / best / 1e9}
        if name.startswith("symmetric"):
            res[name]["symmetry_error"] = (out[:2048, 30000:32048] - out[30000:32048, :2048].T).abs().max().item()
        print(name, res[name], flush=True)
json.dump(res, open("gpurun_out/c5_full_gram.json", "w"), indent=1)
