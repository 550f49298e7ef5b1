"""Multi-GPU timings of BASELINE configs C4 and C5 (run under torchrun, NCCL): per rank the feature maps of its x rows
and of all y, its row block of the Gram (no collective), and the phase- / feature-sharded prior sample (one all-reduce of
N doubles).  Times are CUDA-event times, max over ranks; rank 0 prints one JSON object.
    python -m torch.distributed.run --nproc-per-node G tools/config_bench_multi.py [--gram-rows-cap R]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from liestationarykernels_b200 import distributed as D
from liestationarykernels_b200 import ops
from liestationarykernels_b200.prior_approximation import RandomFourierApproximation, RandomPhaseApproximation
from liestationarykernels_b200.spaces import HyperbolicSpace, Stiefel, SymmetricPositiveDefiniteMatrices
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomFourierFeatureKernel, RandomPhaseKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


res = {"world": world}
with torch.no_grad():
    torch.manual_seed(0)                      # every rank draws the same points, phases and weights
    # ---- C4: Stiefel V(5,3), N = 32768, P = 2048 ----
    N, P = 32768, 2048
    sp = Stiefel(5, 3, order=10, average_order=30)
    meas = MaternSpectralMeasure(9, 1.0, 2.5)
    rk = RandomPhaseKernel(meas, sp, phase_order=P); rk.eval()
    x = sp.rand(N)
    lo, hi = D.row_shard(N, world, rank)
    xl = x[lo:hi].contiguous()
    t_feat = timeit(lambda: (rk._features(xl, rk._row_scales(), panel=True), rk._features(x, rk._row_scales(), panel=True)), 2, 1)
    phi_l, phi = rk._features(xl, rk._row_scales(), panel=True), rk._features(x, rk._row_scales(), panel=True)
    out = torch.empty(hi - lo, N, dtype=torch.float64, device="cuda")
    t_gram = timeit(lambda: ops.gram(phi_l, phi, 1.0, out=out), 2, 1)
    es = EigenbasisSumKernel(meas, sp); es.eval()
    samp = D.PhaseShardedSampler(RandomPhaseApproximation(es, phase_order=P))
    t_samp = timeit(lambda: samp(x), 3, 1)
    res["C4_stiefel53"] = dict(N=N, P=P, features_local_plus_all_y_ms=t_feat, gram_row_block_ms=t_gram,
                               gram_tflops_aggregate=2.0 * N * N * phi.kg * 4 / t_gram / 1e9, sharded_sampler_ms=t_samp)
    del out, phi, phi_l
    torch.cuda.empty_cache()
    # ---- C5: H^3 and SPD(3), N = 65536, m = 8192 ----
    N, m = 65536, 8192
    for name, sp, meas in (("C5_h3_matern", HyperbolicSpace(3, order=m), MaternSpectralMeasure(3, 1.0, 2.5)),
                           ("C5_spd3_matern", SymmetricPositiveDefiniteMatrices(3, order=m), MaternSpectralMeasure(6, 1.0, 2.5))):
        fk = RandomFourierFeatureKernel(meas, sp); fk.eval()
        x = sp.rand(N)
        lo, hi = D.row_shard(N, world, rank)
        rows = hi - lo
        if world == 1:
            rows = N // 8                      # one GPU cannot hold the 34 GB Gram next to the features: time an eighth
        xg = sp.to_group(x)
        xl = xg[lo:lo + rows].contiguous()
        t_feat = timeit(lambda: (sp._features(xl, sp.lb_eigenspaces, 1.0, "panel"), sp._features(xg, sp.lb_eigenspaces, 1.0, "panel")), 2, 1)
        psi_l, psi = sp._features(xl, sp.lb_eigenspaces, 1.0, "panel"), sp._features(xg, sp.lb_eigenspaces, 1.0, "panel")
        out = torch.empty(rows, N, dtype=torch.float64, device="cuda")
        t_gram = timeit(lambda: ops.gram(psi_l, psi, 1.0, out=out), 2, 1)
        samp = D.FeatureShardedSampler(RandomFourierApproximation(fk))
        t_samp = timeit(lambda: samp(x), 3, 1)
        res[name] = dict(N=N, m=m, rows_per_rank=rows, features_local_plus_all_y_ms=t_feat, gram_row_block_ms=t_gram,
                         gram_tflops_aggregate=2.0 * rows * world * N * 2 * m / t_gram / 1e9, sharded_sampler_ms=t_samp)
        del out, psi, psi_l
        torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
