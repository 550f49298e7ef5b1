"""Multi-GPU check (run under torchrun, NCCL): row-sharded covariance reassembles the single-GPU matrix, and the
phase-sharded RandomPhaseApproximation all-reduces to the unsharded sample.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from liestationarykernels_b200 import distributed as D
from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
from liestationarykernels_b200.spaces import SO, Stiefel
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
with torch.no_grad():
    torch.manual_seed(0)                       # same draws on every rank
    so5 = SO(5, order=20)
    kern = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), so5)
    kern.eval()
    x, y = so5.rand(1000), so5.rand(777)
    full = kern(x, y)
    block, lo, hi = D.sharded_covariance(kern, x, y)
    res["rows_max_abs_diff"] = (block - full[lo:hi]).abs().max().item()
    # phase-sharded sampler on a homogeneous space, all-reduce over NCCL
    st = Stiefel(5, 3, order=10, average_order=30)
    es = EigenbasisSumKernel(MaternSpectralMeasure(9, 1.0, 2.5), st)
    es.eval()
    sampler = RandomPhaseApproximation(es, phase_order=256)
    pts = st.rand(4096)
    ref = sampler(pts)
    sharded = D.PhaseShardedSampler(sampler)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    got = sharded(pts)
    e0.record()
    for _ in range(5):
        got = sharded(pts)
    e1.record()
    torch.cuda.synchronize()
    res["sample_rel_diff"] = ((got - ref).abs().max() / ref.abs().max()).item()
    res["sharded_sampler_ms"] = e0.elapsed_time(e1) / 5
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        sampler(pts)
    t1.record()
    torch.cuda.synchronize()
    res["single_gpu_sampler_ms"] = t0.elapsed_time(t1) / 5
    # feature-sharded sampler on the non-compact spaces (RandomFourierApproximation), all-reduce over NCCL
    from liestationarykernels_b200.prior_approximation import RandomFourierApproximation
    from liestationarykernels_b200.spaces import HyperbolicSpace, SymmetricPositiveDefiniteMatrices
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel
    for tag, space, meas in (("h3", HyperbolicSpace(3, order=1000), MaternSpectralMeasure(3, 1.0, 2.5)),
                             ("spd3", SymmetricPositiveDefiniteMatrices(3, order=1000), MaternSpectralMeasure(6, 1.0, 2.5))):
        fk = RandomFourierFeatureKernel(meas, space)
        fk.eval()
        fs = RandomFourierApproximation(fk)
        pts = space.rand(2048)
        ref = fs(pts)
        got = D.FeatureShardedSampler(fs)(pts)
        res["feature_sharded_%s_rel_diff" % tag] = ((got - ref).abs().max() / ref.abs().max()).item()
res["world"] = world
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    print(json.dumps({"ranks": gathered}))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"ranks": gathered}, open("gpurun_out/dist_check.json", "w"), indent=1)
dist.destroy_process_group()
