"""Development / measurement: lsk_potrf_upper + lsk_potrs_upper against LAPACK (CPU, parity) and cuSOLVER (GPU, speed)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.linalg import CholeskyFactor

torch.manual_seed(0)
dev = "cuda"
out = {"parity": [], "speed": []}


def spd(n, cond_shift=1.0):
    g = torch.randn(n, n + 8, dtype=torch.float64)
    return g @ g.T / (n + 8) + cond_shift * torch.eye(n, dtype=torch.float64)


for n in [1, 2, 5, 64, 127, 128, 129, 255, 256, 257, 300, 1000, 1537, 2085]:
    a = spd(n, 0.05)
    ref = torch.linalg.cholesky(a)
    f = CholeskyFactor(a.to(dev))
    L = f.L.cpu()
    err = ((L - ref).abs().max() / ref.abs().max()).item()
    b = torch.randn(n, dtype=torch.float64)
    x_ref = torch.cholesky_solve(b[:, None], ref)[:, 0]
    x = f.solve(b.to(dev)).cpu()
    y = f.solve_lower(b.to(dev)).cpu()
    y_ref = torch.linalg.solve_triangular(ref, b[:, None], upper=False)[:, 0]
    serr = ((x - x_ref).abs().max() / x_ref.abs().max()).item()
    yerr = ((y - y_ref).abs().max() / y_ref.abs().max()).item()
    lerr = abs(f.logdet().item() - 2 * torch.log(torch.diagonal(ref)).sum().item())
    inv_ref = torch.cholesky_inverse(ref)
    inv = f.inverse().cpu()
    ierr = ((inv - inv_ref).abs().max() / inv_ref.abs().max()).item()
    assert torch.equal(inv, inv.T)
    out["parity"].append({"n": n, "L_rel": err, "solve_rel": serr, "fwd_rel": yerr, "logdet_abs": lerr, "inverse_rel": ierr, "info": f.info})
    print(out["parity"][-1], flush=True)

# not positive definite
a = spd(300, 0.05); a[200, 200] = -1.0
f = CholeskyFactor(a.to(dev), check=False)
print("info for an indefinite matrix (expect 201):", f.info)
out["indefinite_info"] = f.info

for n in [4096, 8192, 16384]:
    g = torch.randn(n, 256, dtype=torch.float64, device=dev)
    a = g @ g.T / 256 + torch.eye(n, dtype=torch.float64, device=dev)
    del g
    work = a.clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for it in range(4):
        work.copy_(a)
        torch.cuda.synchronize()
        ev[0].record()
        f = CholeskyFactor(work, overwrite=True, check=False)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    info = f.info
    b = torch.randn(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ev[0].record(); x = f.solve(b); ev[1].record(); torch.cuda.synchronize()
    t_solve = ev[0].elapsed_time(ev[1])
    resid = ((a @ x - b).norm() / b.norm()).item()
    lib_best = 1e9
    for it in range(3):
        torch.cuda.synchronize()
        ev[0].record(); Lc = torch.linalg.cholesky(a); ev[1].record(); torch.cuda.synchronize()
        lib_best = min(lib_best, ev[0].elapsed_time(ev[1]))
    ev[0].record(); xc = torch.cholesky_solve(b[:, None], Lc); ev[1].record(); torch.cuda.synchronize()
    t_lib_solve = ev[0].elapsed_time(ev[1])
    diff = ((torch.triu(f.matrix).mT - Lc).abs().max() / Lc.abs().max()).item()
    t_inv = 1e9
    for it in range(2):
        torch.cuda.synchronize()
        ev[0].record(); inv = f.inverse(); ev[1].record(); torch.cuda.synchronize()
        t_inv = min(t_inv, ev[0].elapsed_time(ev[1]))
    ev[0].record(); inv_lib = torch.cholesky_inverse(Lc); ev[1].record(); torch.cuda.synchronize()
    t_inv_lib = ev[0].elapsed_time(ev[1])
    inv_diff = ((inv - inv_lib).abs().max() / inv_lib.abs().max()).item()
    inv_resid = ((a @ inv[:, :64] - torch.eye(n, 64, dtype=torch.float64, device=dev)).abs().max()).item()
    del inv, inv_lib
    rec = {"n": n, "potrf_ms": best, "tflops": n ** 3 / 3 / best / 1e9, "cusolver_potrf_ms": lib_best,
           "cusolver_tflops": n ** 3 / 3 / lib_best / 1e9, "potrs_ms": t_solve, "cusolver_potrs_ms": t_lib_solve,
           "residual": resid, "L_vs_cusolver": diff, "potri_ms": t_inv, "potri_tflops": 2 * n ** 3 / 3 / t_inv / 1e9,
           "cusolver_potri_ms": t_inv_lib, "inverse_vs_cusolver": inv_diff, "inverse_residual": inv_resid, "info": info}
    out["speed"].append(rec)
    print(rec, flush=True)
    del a, work, Lc, f
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/chol_bench.json", "w"), indent=1)
