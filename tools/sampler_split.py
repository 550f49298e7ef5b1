"""Development (torchrun, 2+ ranks): where the phase-sharded C4 sampler spends its time - full unsharded sample, the local shard alone,
the all-reduce alone, and the sharded call."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from liestationarykernels_b200 import distributed as D
from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
from liestationarykernels_b200.spaces import Stiefel
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def timeit(fn, reps=5, warm=2):
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
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
N, P = 32768, 2048
sp = Stiefel(5, 3, order=10, average_order=30)
es = EigenbasisSumKernel(MaternSpectralMeasure(9, 1.0, 2.5), sp).eval()
full = RandomPhaseApproximation(es, phase_order=P)
sh = D.PhaseShardedSampler(full)
x = sp.rand(N).contiguous()
buf = torch.zeros(N, dtype=torch.float64, device="cuda")
res = {"world": world}
with torch.no_grad():
    res["full_ms"] = timeit(lambda: full(x))
    import copy
    loc = copy.copy(full)
    loc.phases = full.phases[sh.lo:sh.hi].contiguous()
    loc.weights = full.weights[sh.weight_index].contiguous()
    loc.phase_order = sh.hi - sh.lo
    res["local_shard_ms"] = timeit(lambda: loc(x))
    if world > 1:
        res["allreduce_ms"] = timeit(lambda: dist.all_reduce(buf))
    res["sharded_ms"] = timeit(lambda: sh(x))
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
