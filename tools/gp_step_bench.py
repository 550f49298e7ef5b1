"""Measurement: one training step of the reference's GP regression loop (examples/gpr_model.py:59-66: loss = -mll, backward,
Adam step) on SO(5) / Matern-5/2 / 100 irreps at n training points, with the time of its parts."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200 import gpr
from liestationarykernels_b200.linalg import CholeskyFactor
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure


def timed(fn, reps=3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        ev[0].record()
        out = fn()
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return best, out


torch.manual_seed(0)
space = SO(5, order=100)
res = []
for n in (4096, 8192, 16384):
    meas = MaternSpectralMeasure(10, 1.0, 2.5)
    kern = EigenbasisSumKernel(meas, space)
    lik = gpr.GaussianLikelihood(device="cuda")
    x = space.rand(n)
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    flat = x.reshape(n, -1)
    model = gpr.ExactGPModel(flat, y, lik, kern, space, point_shape=(5, 5))
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    mll = gpr.ExactMarginalLogLikelihood(lik, model)

    def step():
        opt.zero_grad()
        loss = -mll(model(flat), y)
        loss.backward()
        opt.step()
        return loss

    step()
    t_step, loss = timed(step)
    with torch.no_grad():
        kern.eval()
        t_kernel, k = timed(lambda: kern(x))
        a = k + (lik.noise.detach() + 1e-6) * torch.eye(n, dtype=torch.float64, device="cuda")
        t_potrf, f = timed(lambda: CholeskyFactor(a, overwrite=False, check=False))
        t_potrs, _ = timed(lambda: f.solve(y))
        t_potri, _ = timed(lambda: f.inverse(), reps=2)
    rec = {"n": n, "training_step_ms": t_step, "loss": loss.item(), "kernel_symmetric_ms": t_kernel, "potrf_ms_incl_copy": t_potrf,
           "potrs_ms": t_potrs, "potri_ms": t_potri,
           "lengthscale_grad": meas.lengthscale.grad.item(), "noise_grad": lik.raw_noise.grad.item()}
    res.append(rec)
    print(rec, flush=True)
    del model, a, f, k
    torch.cuda.empty_cache()
json.dump(res, open("gpurun_out/gp_step_bench.json", "w"), indent=1)
