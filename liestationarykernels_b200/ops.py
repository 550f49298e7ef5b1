"""Thin torch-facing wrappers over the C ABI: tensors in, tensors out, launches on the current stream."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .classfn import GroupPlan

F64 = torch.float64


def _as_matrix_batch(x: torch.Tensor, n: int, is_complex: bool) -> torch.Tensor:
    want = torch.complex128 if is_complex else F64
    if x.dim() != 3 or x.shape[-1] != n or x.shape[-2] != n:
        raise ValueError("expected a [num, %d, %d] batch of group elements, got %s" % (n, n, tuple(x.shape)))
    if not x.is_cuda:
        raise _lib.LskError("inputs must live on the CUDA device (no CPU fallback)")
    if x.dtype != want:
        x = x.to(want)
    return x.contiguous()        # G8: Stiefel.rand-style strided views are made contiguous before pointers are taken


def embed(plan: GroupPlan, x: torch.Tensor, side: str) -> torch.Tensor:
    """Compound-matrix embedding of a batch of group elements in the panel-major operand layout
    (replaces `inv` + `cartesian_prod` operand preparation, space.py:62-78)."""
    _lib.require_cuda()
    lib = _lib.load()
    x = _as_matrix_batch(x, plan.n, plan.is_complex)
    num = x.shape[0]
    segs = plan.segments_a if side == "a" else plan.segments_b
    nseg = len(segs)
    wedge = (ctypes.c_int * nseg)(*[s[0] for s in segs])
    part = (ctypes.c_int * nseg)(*[s[1] for s in segs])
    begin = (ctypes.c_int * nseg)(*plan.group_begin)
    size = lib.lsk_embed_buffer_doubles(num, plan.kg_total)
    out = torch.empty(max(size, 1), dtype=F64, device=x.device)
    if num:
        xr = torch.view_as_real(x) if plan.is_complex else x
        _lib.check(lib.lsk_embed_points(_lib.ptr(xr), num, plan.n, int(plan.is_complex), nseg, wedge, part, begin,
                                        plan.kg_total, _lib.ptr(out), _lib.stream_ptr()), "lsk_embed_points")
        _lib.count_launch()
    return out


def pairwise_poly(emb_a: torch.Tensor, emb_b: torch.Tensor, n_rows: int, n_cols: int, kg_total: int,
                  epilogue: _lib.LskEpilogue, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[i, j] = scale * P(invariants of pair (i, j)) — the fused covariance tile kernel."""
    _lib.require_cuda()
    lib = _lib.load()
    if out is None:
        out = torch.empty((n_rows, n_cols), dtype=F64, device=emb_a.device)
    else:
        if out.dtype != F64 or out.dim() != 2 or out.shape[0] != n_rows or out.shape[1] != n_cols or out.stride(1) != 1:
            raise ValueError("out must be a float64 [n_rows, n_cols] tensor with unit column stride")
    if n_rows and n_cols:
        _lib.check(lib.lsk_pairwise_poly(_lib.ptr(emb_a), _lib.ptr(emb_b), n_rows, n_cols, kg_total,
                                         ctypes.byref(epilogue), _lib.ptr(out), out.stride(0), _lib.stream_ptr()),
                   "lsk_pairwise_poly")
        _lib.count_launch()
    return out
