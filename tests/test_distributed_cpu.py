"""N > 1 host logic on CPU (gloo, world_size 2): row sharding reassembles the full covariance, and the phase-sharded
sampler (weights laid out irrep-major / phase-minor) all-reduces to the unsharded sample.  The per-rank compute
stand-in is the oracle port, so these tests run without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lsk_oracle as orc
    from liestationarykernels_b200 import distributed as D
    torch.manual_seed(0)
    space = orc.Group("SO", 3, 6)
    meas = orc.Measure("heat", 3, 1.0)
    x, y = space.rand(11), space.rand(7)
    # ---- row-sharded covariance: gather of the shards equals the full matrix
    lo, hi = D.row_shard(x.shape[0], world, rank)
    local = orc.eigensum_unnormalised(space, meas, x[lo:hi], y)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, local))
    full = torch.cat([p[2] for p in sorted(parts, key=lambda t: t[0])])
    ref = orc.eigensum_unnormalised(space, meas, x, y)
    ok_rows = torch.allclose(full, ref, rtol=1e-13, atol=0) and full.shape == ref.shape
    # ---- phase-sharded sampler
    P, L = 10, len(space.irreps)
    phases = space.rand(P)
    weights = torch.randn(P * L, dtype=torch.float64)
    norm = torch.tensor(2.5, dtype=torch.float64)
    ref_sample = orc.random_phase_sample(space, meas, phases, weights, x, norm)
    plo, phi, idx = D.phase_shard(P, L, world, rank)
    part = orc.random_phase_sample(space, meas, phases[plo:phi], weights[idx], x, norm) * ((phi - plo) / P) ** 0.5
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    ok_sample = torch.allclose(part, ref_sample, rtol=1e-12, atol=1e-14)
    ret[rank] = (bool(ok_rows), bool(ok_sample), (lo, hi), (plo, phi))
    dist.destroy_process_group()


def test_row_and_phase_sharding_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        ok_rows, ok_sample, _, _ = ret[rank]
        assert ok_rows and ok_sample
    assert ret[0][2][1] == ret[1][2][0] and ret[0][3][1] == ret[1][3][0]


def test_shard_bounds_cover_everything():
    from liestationarykernels_b200.distributed import phase_shard, row_shard
    for n in (0, 1, 7, 16384):
        for w in (1, 2, 3, 8):
            cuts = [row_shard(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
    seen = torch.cat([phase_shard(10, 4, 3, r)[2] for r in range(3)])
    assert sorted(seen.tolist()) == list(range(40))
