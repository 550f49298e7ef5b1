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


def _embed_doubles(num: int, kg_total: int) -> int:
    """= lsk_embed_buffer_doubles(num, kg_total): rows padded to 128, 256 doubles per (64-row panel, k-group)"""
    return ((num + 127) // 128) * 2 * kg_total * 256


def _pair_args(plan: GroupPlan):
    """ctypes argument arrays of lsk_embed_points2 for a plan, built once (the hot call is launch-latency bound at small sizes)"""
    cached = getattr(plan, "_pair_args", None)
    if cached is None:
        nseg = len(plan.segments_a)
        cached = (nseg, (ctypes.c_int * nseg)(*[s[0] for s in plan.segments_a]), (ctypes.c_int * nseg)(*[s[1] for s in plan.segments_a]),
                  (ctypes.c_int * nseg)(*[s[1] for s in plan.segments_b]), (ctypes.c_int * nseg)(*plan.group_begin))
        plan._pair_args = cached
    return cached


def embed_pair(plan: GroupPlan, x: torch.Tensor, y: torch.Tensor, snapshot=None):
    """Embeddings of both operands of `kernel(x, y)` with ONE launch (side A for x, side B for y).  `snapshot` =
    (pointer array, count, pinned destination, sequence number): the same launch also writes the hyper-parameter snapshot
    (`lsk_embed_points2_snapshot`); when the one-launch path does not apply it is written by a launch of its own."""
    _lib.require_cuda()
    lib = _lib.load()
    x = _as_matrix_batch(x, plan.n, plan.is_complex)
    y = _as_matrix_batch(y, plan.n, plan.is_complex)
    if x.shape[0] == 0 or y.shape[0] == 0 or [s[0] for s in plan.segments_a] != [s[0] for s in plan.segments_b]:
        if snapshot is not None:
            snapshot_launch(snapshot)
        return embed(plan, x, "a"), embed(plan, y, "b")
    nseg, wedge, part_a, part_b, begin = _pair_args(plan)
    na, nb = _embed_doubles(x.shape[0], plan.kg_total), _embed_doubles(y.shape[0], plan.kg_total)
    both = torch.empty(na + nb, dtype=F64, device=x.device)          # one allocation for the two operands
    ea, eb = both[:na], both[na:]
    xr, yr = (torch.view_as_real(x), torch.view_as_real(y)) if plan.is_complex else (x, y)
    if snapshot is None:
        _lib.check(lib.lsk_embed_points2(_lib.ptr(xr), x.shape[0], part_a, _lib.ptr(ea), _lib.ptr(yr), y.shape[0], part_b,
                                         _lib.ptr(eb), plan.n, int(plan.is_complex), nseg, wedge, begin, plan.kg_total,
                                         _lib.stream_ptr()), "lsk_embed_points2")
    else:
        arr, count, dst, seq = snapshot
        _lib.check(lib.lsk_embed_points2_snapshot(_lib.ptr(xr), x.shape[0], part_a, _lib.ptr(ea), _lib.ptr(yr), y.shape[0], part_b,
                                                  _lib.ptr(eb), plan.n, int(plan.is_complex), nseg, wedge, begin, plan.kg_total,
                                                  arr, count, dst, seq, _lib.stream_ptr()), "lsk_embed_points2_snapshot")
    _lib.count_launch()
    return ea, eb


def snapshot_launch(snapshot):
    """stand-alone `lsk_snapshot_doubles` launch for a (pointer array, count, pinned destination, sequence number) tuple"""
    arr, count, dst, seq = snapshot
    _lib.check(_lib.load().lsk_snapshot_doubles(arr, count, dst, seq, _lib.stream_ptr()), "lsk_snapshot_doubles")


def pairwise_poly(emb_a: torch.Tensor, emb_b: torch.Tensor, n_rows: int, n_cols: int, kg_total: int,
                  epilogue: _lib.LskEpilogue, out: torch.Tensor | None = None, symmetric: bool = False) -> torch.Tensor:
    """out[i, j] = scale * P(invariants of pair (i, j)) — the fused covariance tile kernel.
    symmetric=True (same points on both sides, n_rows == n_cols): only the upper triangle of tiles is computed and
    mirrored, which halves the work of `kernel(x)` and of a Gram `Phi Phi^T`."""
    _lib.require_cuda()
    lib = _lib.load()
    if out is None:
        out = torch.empty((n_rows, n_cols), dtype=F64, device=emb_a.device)
    else:
        if out.dtype != F64 or out.dim() != 2 or out.shape[0] != n_rows or out.shape[1] != n_cols or out.stride(1) != 1:
            raise ValueError("out must be a float64 [n_rows, n_cols] tensor with unit column stride")
    reserved = epilogue.reserved                      # the struct may be a cached one: flag set for this launch only
    epilogue.reserved = (reserved | 2) if (symmetric and n_rows == n_cols and n_rows > 256) else (reserved & ~2)
    try:
        if n_rows and n_cols:
            _lib.check(lib.lsk_pairwise_poly(_lib.ptr(emb_a), _lib.ptr(emb_b), n_rows, n_cols, kg_total,
                                             ctypes.byref(epilogue), _lib.ptr(out), out.stride(0), _lib.stream_ptr()),
                       "lsk_pairwise_poly")
            _lib.count_launch()
    finally:
        epilogue.reserved = reserved
    return out


# ----------------------------------------------------------------------------------------------------------
# panel-major operands (the layout the pairwise kernel reads and, for feature maps, writes)
# ----------------------------------------------------------------------------------------------------------
class PanelMatrix:
    """A [rows, cols] float64 matrix stored as [row / 64][col / 4][row % 64][col % 4] (rows zero-padded to a
    multiple of 128, cols to a multiple of 4): one pipeline stage of a 64-row panel is contiguous, and an fp64
    tensor-core fragment is a conflict-free 256-byte read."""

    def __init__(self, rows: int, cols: int, device, buf: torch.Tensor | None = None):
        self.rows, self.cols = rows, cols
        self.kg = max((cols + 3) // 4, 1)
        self.row_panels = ((rows + 127) // 128) * 2
        n = self.row_panels * self.kg * 256
        self.buf = torch.zeros(max(n, 1), dtype=F64, device=device) if buf is None else buf

    @staticmethod
    def for_full_write(rows: int, cols: int, device) -> "PanelMatrix":
        """Target of a kernel that writes every (row < rows, col < cols) entry: zero-filled only if the layout has padding
        rows / columns (a feature map of C5 is 8.6 GB: the fill would cost as much as writing it)."""
        if rows % 128 == 0 and cols % 4 == 0 and rows and cols:
            kg, panels = cols // 4, rows // 64
            return PanelMatrix(rows, cols, device, buf=torch.empty(panels * kg * 256, dtype=F64, device=device))
        return PanelMatrix(rows, cols, device)

    @staticmethod
    def from_dense(dense: torch.Tensor) -> "PanelMatrix":
        rows, cols = dense.shape
        pm = PanelMatrix(rows, cols, dense.device)
        if rows and cols:
            padded = torch.zeros((pm.row_panels * 64, pm.kg * 4), dtype=F64, device=dense.device)
            padded[:rows, :cols] = dense
            pm.buf = padded.view(pm.row_panels, 64, pm.kg, 4).permute(0, 2, 1, 3).contiguous().view(-1)
        return pm

    def row_block(self, r0: int, r1: int) -> "PanelMatrix":
        """Rows [r0, r1) as a PanelMatrix sharing this buffer (no copy): a 128-row-aligned block of the panel-major
        layout is contiguous.  Used by the multi-GPU Gram to address row / column blocks of one feature map."""
        if r0 % 128 or not (0 <= r0 <= r1 <= self.rows):
            raise ValueError("row block must start on a multiple of 128 inside the matrix")
        pm = PanelMatrix.__new__(PanelMatrix)
        pm.rows, pm.cols, pm.kg = r1 - r0, self.cols, self.kg
        pm.row_panels = ((pm.rows + 127) // 128) * 2
        pm.buf = self.buf[(r0 // 64) * self.kg * 256:]
        return pm

    def to_dense(self) -> torch.Tensor:
        v = self.buf[: self.row_panels * self.kg * 256].view(self.row_panels, self.kg, 64, 4).permute(0, 2, 1, 3)
        return v.reshape(self.row_panels * 64, self.kg * 4)[: self.rows, : self.cols].contiguous()


def gram(a: PanelMatrix, b: PanelMatrix, scale: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """scale * A B^T on the fp64 tensor-core path (the Gram the reference's lazy `MatmulLazyTensor` evaluates to,
    spectral_kernel.py:125-127, 195-197)."""
    if a.kg != b.kg:
        raise ValueError("operands have different inner dimensions")
    ep = _lib.LskEpilogue()
    ep.kind, ep.nforms, ep.nvars, ep.nterms = _lib.EPI_LINEAR, 1, 1, 1
    ep.group_begin[0], ep.group_end[0] = 0, a.kg
    ep.scale = float(scale)
    ep.out_layout = _lib.OUT_ROW_MAJOR
    return pairwise_poly(a.buf, b.buf, a.rows, b.rows, a.kg, ep, out=out, symmetric=(a is b))


def gram_columns(a: PanelMatrix, b: PanelMatrix, g0: int, g1: int, scale: float = 1.0) -> torch.Tensor:
    """scale * A[:, 4 g0 : 4 g1] B[:, 4 g0 : 4 g1]^T: the Gram of one column block (k-groups g0 .. g1 - 1) of two feature maps,
    read in place — the operand pointers are advanced by g0 k-groups and the full row stride is kept."""
    if a.kg != b.kg or not (0 <= g0 < g1 <= a.kg):
        raise ValueError("bad column block")
    _lib.require_cuda()
    lib = _lib.load()
    ep = _lib.LskEpilogue()
    ep.kind, ep.nforms, ep.nvars, ep.nterms = _lib.EPI_LINEAR, 1, 1, 1
    ep.group_begin[0], ep.group_end[0] = 0, g1 - g0
    ep.scale = float(scale)
    ep.out_layout = _lib.OUT_ROW_MAJOR
    out = torch.empty((a.rows, b.rows), dtype=F64, device=a.buf.device)
    if a.rows and b.rows:
        _lib.check(lib.lsk_pairwise_poly(ctypes.c_void_p(a.buf.data_ptr() + g0 * 256 * 8), ctypes.c_void_p(b.buf.data_ptr() + g0 * 256 * 8),
                                         a.rows, b.rows, a.kg, ctypes.byref(ep), _lib.ptr(out), out.stride(0), _lib.stream_ptr()),
                   "lsk_pairwise_poly(column block)")
        _lib.count_launch()
    return out


def panel_matvec(a: PanelMatrix, w: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * A w for a panel-major A [rows, cols] and a dense vector w [cols]: the prior sample `einsum('nm,m->n')` of
    prior_approximation.py:85-88 as one HBM-bound pass over the feature map (`lsk_panel_gemv`)."""
    _lib.require_cuda()
    lib = _lib.load()
    if w.numel() != a.cols:
        raise ValueError("weight vector has %d entries, the feature map %d columns" % (w.numel(), a.cols))
    wp = torch.zeros(a.kg * 4, dtype=F64, device=a.buf.device)
    wp[: a.cols] = w.to(F64).reshape(-1)
    out = torch.empty(a.rows, dtype=F64, device=a.buf.device)
    if a.rows:
        _lib.check(lib.lsk_panel_gemv(_lib.ptr(a.buf), a.rows, a.kg, _lib.ptr(wp), float(scale), _lib.ptr(out),
                                      _lib.stream_ptr()), "lsk_panel_gemv")
        _lib.count_launch()
    return out


def class_features(plan: GroupPlan, x: torch.Tensor, phases: torch.Tensor, row_scales, panel: bool):
    """Per-irrep class-function features  Phi[n, l * P + p] = row_scales[l] * Re chi_l(x_n phase_p^{-1})
    (`make_embedding`, spectral_kernel.py:96-109 / prior_approximation.py:70-83: irrep-major, phase-minor columns).
    Returns a PanelMatrix (panel=True, ready to be a Gram operand) or a dense [N, L*P] tensor."""
    _lib.require_cuda()
    N, P, L = x.shape[0], phases.shape[0], len(plan.irreps)
    if panel and P % 4:
        # irrep column blocks must start on a k-group boundary in the panel layout: go through the dense form
        return PanelMatrix.from_dense(class_features(plan, x, phases, row_scales, panel=False))
    return features_from_operands(plan, embed(plan, x, "a"), embed(plan, phases, "b"), N, P, row_scales, panel, x.device)


def features_from_operands(plan, ea: torch.Tensor, eb: torch.Tensor, N: int, P: int, row_scales, panel: bool, device):
    """The per-irrep feature launches of `class_features` on ready panel-major operands (`plan` supplies the bilinear
    forms and the per-irrep polynomials: a classfn.GroupPlan or the zonal plan of spaces/sphere.py)."""
    L = len(plan.irreps)
    if panel and P % 4:
        dense = features_from_operands(plan, ea, eb, N, P, row_scales, False, device)
        return PanelMatrix.from_dense(dense)
    if panel:
        target = PanelMatrix.for_full_write(N, L * P, device)
        out = target.buf
    else:
        target = torch.zeros((N, L * P), dtype=F64, device=device)
        out = target
    for ep, col_shift in plan.feature_epilogues(row_scales, P):
        if panel:
            ep.out_layout, ep.out_kg = _lib.OUT_PANEL_MAJOR, target.kg
            view = out
            ld = 0
        else:
            ep.out_layout = _lib.OUT_ROW_MAJOR
            view = out
            ld = out.stride(0)
        lib = _lib.load()
        # column block of this irrep batch starts at col_shift: shift the base pointer (row-major) or the k-group
        # origin (panel-major; col_shift is a multiple of 4 there because it is l * P and P % 4 == 0 is required)
        if panel:
            if col_shift % 4:
                raise ValueError("panel-major features need the number of phases to be a multiple of 4")
            base = view.data_ptr() + (col_shift // 4) * 256 * 8
        else:
            base = view.data_ptr() + col_shift * 8
        if N and P:
            _lib.check(lib.lsk_pairwise_poly(_lib.ptr(ea), _lib.ptr(eb), N, P, plan.kg_total, ctypes.byref(ep),
                                             ctypes.c_void_p(base), ld, _lib.stream_ptr()), "lsk_pairwise_poly(features)")
            _lib.count_launch()
    return target
