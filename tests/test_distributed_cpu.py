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


def test_triangular_chunks_cover_the_rows_and_balance_the_upper_triangle():
    """`triangular_chunks`: 2W chunks, rank r owns r and 2W-1-r; together they tile [0, n) exactly, boundaries sit on multiples of
    128 (panel-major row blocks must), and the tile work of the upper triangle is balanced at BASELINE sizes."""
    from liestationarykernels_b200.distributed import triangular_chunks
    for n in (32768, 65536, 16384, 5000, 130, 1):
        for w in (1, 2, 3, 4, 8):
            per_rank = [triangular_chunks(n, w, r) for r in range(w)]
            chunks = sorted(c for cs in per_rank for c in cs)
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
            assert all(lo % 128 == 0 for lo, _ in chunks)
            if n >= 16384:
                work = [sum((hi - lo) * (n - lo) - (hi - lo) ** 2 / 2 for lo, hi in cs) for cs in per_rank]
                assert max(work) / min(work) < 1.01


def _tri_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from liestationarykernels_b200 import distributed as D
    torch.manual_seed(0)
    n, k = 700, 24
    phi = torch.randn(n, k, dtype=torch.float64)
    full = phi @ phi.T
    ok = True
    covered = torch.zeros(n, n, dtype=torch.bool)
    for lo, hi in D.triangular_chunks(n, world, rank):
        block = phi[lo:hi] @ phi[lo:].T                    # the per-chunk product `sharded_symmetric_gram` launches
        ok = ok and torch.allclose(block, full[lo:hi, lo:], rtol=1e-13, atol=1e-13)
        covered[lo:hi, lo:] = True
    parts = [None] * world
    dist.all_gather_object(parts, covered)
    union = torch.stack(parts).any(dim=0)
    ret[rank] = (bool(ok), bool(torch.equal(union | union.T, torch.ones(n, n, dtype=torch.bool))), int(sum(p.sum() for p in parts)))
    dist.destroy_process_group()


def test_triangular_sharding_world2_covers_the_upper_triangle_once():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_tri_worker, args=(2, port, ret), nprocs=2, join=True)
    assert all(ret[r][0] and ret[r][1] for r in range(2))
    n = 700
    # every entry on or above the diagonal is owned by exactly one rank (plus the strictly-lower parts of the diagonal blocks)
    assert ret[0][2] >= n * (n + 1) // 2 and ret[0][2] < n * (n + 1) // 2 + 4 * 256 * 256
