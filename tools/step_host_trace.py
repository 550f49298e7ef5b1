"""Development (torchrun): host-side timeline of the C2 step per rank - time from entering kernel(x, y) to the last launch, and the
time spent polling for the hyper-parameter snapshot.  Shows whether the host stays ahead of the GPU (poll wait > 0) or lags it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group(os.environ.get("LSK_BACKEND", "nccl"), device_id=torch.device("cuda", local) if os.environ.get("LSK_BACKEND", "nccl") == "nccl" else None)
torch.manual_seed(0)
space = SO(5, order=100)
kern = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), space).eval()
N = 16384
shard = int(os.environ.get("LSK_SHARD", world))
x_all, y = space.rand(N), space.rand(N)
x = x_all[: N // shard].contiguous()
out = torch.empty(N // shard, N, dtype=torch.float64, device="cuda")
if os.environ.get("LSK_GC_OFF"):
    import gc
    gc.disable()
stamps = []
orig_finish = kern._finish_hyper_read
def finish(p):
    t0 = time.perf_counter()
    r = orig_finish(p)
    stamps.append((t0, time.perf_counter()))
    return r
kern._finish_hyper_read = finish
with torch.no_grad():
    for _ in range(5):
        kern(x, y, out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stamps.clear()
    starts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        starts.append(time.perf_counter())
        kern(x, y, out=out)
    e1.record()
    torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 50
launch_us = [1e6 * (s[0] - st) for s, st in zip(stamps, starts)]
poll_us = [1e6 * (s[1] - s[0]) for s in stamps]
gap_us = [1e6 * (starts[i + 1] - stamps[i][1]) for i in range(len(stamps) - 1)]
res = dict(rank=rank, gpu_ms_per_step=round(gpu_ms, 5), launch_us_median=sorted(launch_us)[25], launch_us_max=max(launch_us),
           poll_us_median=sorted(poll_us)[25], poll_us_min=min(poll_us), zero_polls=sum(1 for p in poll_us if p < 3.0),
           between_calls_us_median=sorted(gap_us)[24])
allres = [None] * world
if world > 1:
    dist.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    for r in allres:
        print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
