"""Multi-GPU orchestration: one process per GPU (`torch.distributed`, NCCL over NVLink on the box, gloo in the CPU
tests).  The reference has no distributed code; the path shards in two ways:

* covariance matrices (EigenbasisSumKernel, Gram of feature maps): x rows are split across ranks, y and the
  coefficient tables are replicated (a few MB), no collective on the data path; the result stays row-sharded.
* prior samples (RandomPhaseApproximation / RandomFourierApproximation): the sum over phases / spherical-function
  features is split across ranks; each rank computes a partial f[N] and one all-reduce(sum) of N doubles finishes it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def row_shard(n: int, world_size: int, rank: int):
    """[lo, hi) of the rows owned by `rank`: contiguous, balanced to within one row, covers [0, n) exactly."""
    return rank * n // world_size, (rank + 1) * n // world_size


def phase_shard(num_phases: int, num_irreps: int, world_size: int, rank: int):
    """Phase range [lo, hi) of `rank` and the indices of its entries in the reference's weight vector, whose layout is
    irrep-major / phase-minor: weights[l * P + p] (prior_approximation.py:61,87)."""
    lo, hi = row_shard(num_phases, world_size, rank)
    idx = (torch.arange(num_irreps)[:, None] * num_phases + torch.arange(lo, hi)[None, :]).reshape(-1)
    return lo, hi, idx


def triangular_chunks(n: int, world_size: int, rank: int, align: int = 128):
    """Balanced ownership of the upper triangle of an n x n symmetric matrix: the rows are cut into 2 W chunks (boundaries on
    multiples of `align`), rank r owns chunks r and 2W-1-r.  Chunk c needs only the columns from its own first row on, so the
    two chunks of a rank together see 2W+1 column chunks on every rank.  Returns [(lo, hi), ...] (empty chunks dropped)."""
    chunks = 2 * world_size
    per = -(-n // chunks)
    per = -(-per // align) * align
    bounds = [min(i * per, n) for i in range(chunks + 1)]
    own = (rank, chunks - 1 - rank)
    return [(bounds[i], bounds[i + 1]) for i in own if bounds[i + 1] > bounds[i]]


def sharded_symmetric_gram(phi, scale: float = 1.0, buffers=None):
    """Upper triangle of scale * Phi Phi^T, row-sharded by `triangular_chunks`, with no collective: every rank holds the whole
    panel-major feature map `phi` (an ops.PanelMatrix; recomputing it costs 1-2 % of the Gram) and evaluates, for each of its
    chunks [lo, hi), the diagonal block on the triangular tile schedule (computed once, mirrored inside the block) and the
    rectangle right of it in general mode.  Returns [(lo, hi, block)], block = float64 [hi-lo, n-lo] holding columns lo..n-1.
    With one rank this is the single symmetric-mode launch of `ops.gram(phi, phi)`."""
    from . import ops
    ws, rk = world()
    n = phi.rows
    if ws == 1:
        out = None if buffers is None else buffers[0]
        return [(0, n, ops.gram(phi, phi, scale, out=out))]
    res = []
    for i, (lo, hi) in enumerate(triangular_chunks(n, ws, rk)):
        out = torch.empty((hi - lo, n - lo), dtype=torch.float64, device=phi.buf.device) if buffers is None else buffers[i]
        rows = phi.row_block(lo, hi)
        ops.gram(rows, rows, scale, out=out[:, : hi - lo])
        if hi < n:
            ops.gram(rows, phi.row_block(hi, n), scale, out=out[:, hi - lo:])
        res.append((lo, hi, out))
    return res


def sharded_covariance(kernel, x, y=None, **kw):
    """Rows [lo, hi) of kernel(x, y) for this rank. Returns (local_block, lo, hi)."""
    ws, rk = world()
    lo, hi = row_shard(x.shape[0], ws, rk)
    if y is None:
        y = x
    return kernel(x[lo:hi].contiguous(), y, **kw), lo, hi


class PhaseShardedSampler:
    """Wraps a RandomPhaseApproximation: every rank keeps a slice of the phases and the matching weights, evaluates the
    partial sample with the fused feature kernels and all-reduces the [N] vector."""

    def __init__(self, sampler):
        ws, rk = world()
        self.full = sampler
        P, L = sampler.phase_order, sampler.approx_order
        lo, hi, idx = phase_shard(P, L, ws, rk)
        self.lo, self.hi = lo, hi
        self.weight_index = idx.to(sampler.weights.device)
        self.world_size = ws

    def __call__(self, x):
        s = self.full
        import copy
        local = copy.copy(s)
        local.phases = s.phases[self.lo:self.hi].contiguous()
        local.weights = s.weights[self.weight_index].contiguous()
        local.phase_order = self.hi - self.lo
        # the 1/sqrt(P) normalisation of the feature map refers to the global number of phases
        part = local(x) * ((self.hi - self.lo) / s.phase_order) ** 0.5 if self.hi > self.lo else torch.zeros(x.shape[0], dtype=torch.float64, device=x.device)
        if self.world_size > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        return part


def feature_shard(num_features: int, world_size: int, rank: int):
    """Slice of the spherical-function features owned by `rank` (RandomFourierApproximation sharding)."""
    return row_shard(num_features, world_size, rank)


class FeatureShardedSampler:
    """Wraps a RandomFourierApproximation (hyperbolic / SPD spaces): every rank keeps a slice of the spherical functions
    (lmd, shift, coeff) and of the two weight vectors, evaluates its partial sample with the fused feature kernel and
    all-reduces the [N] vector (north_star: "the random-phase/feature sum is sharded ... NCCL allreduce")."""

    def __init__(self, sampler):
        ws, rk = world()
        self.full = sampler
        self.lo, self.hi = feature_shard(sampler.kernel.manifold.order, ws, rk)
        self.world_size = ws

    def __call__(self, x):
        from . import ops
        s = self.full
        m = s.kernel.manifold
        funcs = m.lb_eigenspaces
        lo, hi = self.lo, self.hi
        if hi > lo:
            local = type(funcs)(funcs.exp.lmd[lo:hi].contiguous(), funcs.exp.shift[lo:hi].contiguous(), m, funcs.coeff[lo:hi].contiguous())
            w = torch.cat((s.weights_real[lo:hi], -s.weights_imag[lo:hi]))
            part = m._sample(m.to_group(x), local, s._scale(), w)       # feature map fused away (lsk_hyp_sample / lsk_spd_sample)
        else:
            part = torch.zeros(x.shape[0], dtype=torch.float64, device=x.device)
        if self.world_size > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        return part
