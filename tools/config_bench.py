"""Single-GPU timings of the other BASELINE configs (C1, C3, C4, C5) at full size: kernel stages with CUDA events.
Not bench lines (bench.py measures C2); written to gpurun_out/config_bench.json and summarised in profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import ops
from liestationarykernels_b200.prior_approximation import RandomFourierApproximation, RandomPhaseApproximation
from liestationarykernels_b200.spaces import SO, SU, HyperbolicSpace, Stiefel, SymmetricPositiveDefiniteMatrices
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomFourierFeatureKernel, RandomPhaseKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
torch.manual_seed(0)
with torch.no_grad():
    # C1
    sp = SO(3, order=20)
    k = EigenbasisSumKernel(SqExpSpectralMeasure(3, 1.0), sp); k.eval()
    x, y = sp.rand(1024), sp.rand(1024)
    t = timeit(lambda: k(x, y), 20, 3)
    res["C1_so3_heat_1024"] = dict(ms=t, entries_per_s=1024 * 1024 / t * 1e3)
    x, y = sp.rand(16384), sp.rand(16384)
    out = torch.empty(16384, 16384, dtype=torch.float64, device="cuda")
    t = timeit(lambda: k(x, y, out=out), 10, 2)
    res["C1_so3_heat_16384"] = dict(ms=t, entries_per_s=16384 ** 2 / t * 1e3, store_GBs=16384 ** 2 * 8 / t / 1e6)
    # C3
    sp = SU(4, order=20)
    k = EigenbasisSumKernel(SqExpSpectralMeasure(15, 1.0), sp); k.eval()
    x, y = sp.rand(16384), sp.rand(16384)
    t = timeit(lambda: k(x, y, out=out), 5, 2)
    res["C3_su4_heat_16384"] = dict(ms=t, entries_per_s=16384 ** 2 / t * 1e3, tflops_algorithmic_800=16384 ** 2 * 800 / t / 1e9)
    # C2 robust form for comparison
    sp = SO(5, order=100)
    k = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), sp, form="robust"); k.eval()
    x, y = sp.rand(16384), sp.rand(16384)
    t = timeit(lambda: k(x, y, out=out), 5, 2)
    res["C2_so5_robust_form_16384"] = dict(ms=t, entries_per_s=16384 ** 2 / t * 1e3, tflops_algorithmic_756=16384 ** 2 * 756 / t / 1e9)
    del out
    # C4
    N, P = 32768, 2048
    sp = Stiefel(5, 3, order=10, average_order=30)
    meas = MaternSpectralMeasure(9, 1.0, 2.5)
    rk = RandomPhaseKernel(meas, sp, phase_order=P); rk.eval()
    x = sp.rand(N)
    t_feat = timeit(lambda: rk._features(x, rk._row_scales(), panel=True), 3, 1)
    phi = rk._features(x, rk._row_scales(), panel=True)
    gram_out = torch.empty(N, N, dtype=torch.float64, device="cuda")
    t_gram = timeit(lambda: ops.gram(phi, phi, 1.0, out=gram_out), 2, 1)
    flops = 2.0 * N * N * phi.kg * 4
    es = EigenbasisSumKernel(meas, sp); es.eval()
    samp = RandomPhaseApproximation(es, phase_order=P)
    t_samp = timeit(lambda: samp(x), 3, 1)
    res["C4_stiefel53"] = dict(N=N, P=P, A=30, L=10, features_ms=t_feat, pairs_per_s=N * P / t_feat * 1e3,
                               gram_ms=t_gram, gram_tflops=flops / t_gram / 1e9, sampler_ms=t_samp)
    del gram_out, phi
    torch.cuda.empty_cache()
    # C5
    N, m = 65536, 8192
    for name, sp, meas in (("C5_h3_matern", HyperbolicSpace(3, order=m), MaternSpectralMeasure(3, 1.0, 2.5)),
                           ("C5_spd3_matern", SymmetricPositiveDefiniteMatrices(3, order=m), MaternSpectralMeasure(6, 1.0, 2.5))):
        fk = RandomFourierFeatureKernel(meas, sp); fk.eval()
        x = sp.rand(N)
        xg = sp.to_group(x)
        t_feat = timeit(lambda: sp._features(xg, sp.lb_eigenspaces, 1.0, "panel"), 3, 1)
        psi = sp._features(xg, sp.lb_eigenspaces, 1.0, "panel")
        half = N // 4
        gram_out = torch.empty(half, N, dtype=torch.float64, device="cuda")
        sub = ops.PanelMatrix(half, 2 * m, "cuda", buf=psi.buf[: (half // 64) * psi.kg * 256])
        t_gram = timeit(lambda: ops.gram(sub, psi, 1.0, out=gram_out), 2, 1)
        flops = 2.0 * half * N * 2 * m
        samp = RandomFourierApproximation(fk)
        t_samp = timeit(lambda: samp(x), 3, 1)
        res[name] = dict(N=N, m=m, features_ms=t_feat, features_per_s=N * m / t_feat * 1e3, gram_quarter_rows_ms=t_gram,
                         gram_tflops=flops / t_gram / 1e9, gram_full_ms_extrapolated=4 * t_gram, sampler_ms=t_samp)
        del gram_out, psi, sub
        torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/config_bench.json", "w"), indent=1)
