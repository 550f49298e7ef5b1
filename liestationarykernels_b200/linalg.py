"""fp64 Cholesky factorisation and solves on the covariance — the step that follows the kernel in every caller of the
reference (`examples/gpr_model.py:26-39`: `ExactGP.forward` returns `MultivariateNormal(mean, K + 1e-6 I)`, whose
`log_prob` / posterior is a LAPACK / cuSOLVER `potrf` + `potrs`).

Thin torch-facing wrappers over `lsk_potrf_upper` / `lsk_potrs_upper` (csrc/lsk_cholesky.cu, include/lsk.h): blocked
right-looking factorisation whose trailing update runs on the same TMA + DMMA tile kernel as the covariance itself.
No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib

F64 = torch.float64


class NotPositiveDefiniteError(_lib.LskError):
    pass


class CholeskyFactor:
    """A = U^T U.  `matrix` holds U in its upper triangle (row-major; its strict lower triangle is scratch), so the lower
    factor `L = U^T` of `torch.linalg.cholesky` is the same buffer read column-major (`.L` returns that view, masked)."""

    def __init__(self, a: torch.Tensor, overwrite: bool = False, check: bool = True):
        _lib.require_cuda()
        lib = _lib.load()
        if a.dim() != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("expected a square matrix, got %s" % (tuple(a.shape),))
        if not a.is_cuda:
            raise _lib.LskError("inputs must live on the CUDA device (no CPU fallback)")
        if a.dtype != F64:
            raise ValueError("float64 only (the reference computes in float64, spectral_kernel.py:12)")
        if not overwrite or a.stride(1) != 1 or a.stride(0) < a.shape[1]:
            a = a.clone(memory_format=torch.contiguous_format)
        self.matrix = a
        self.n = a.shape[0]
        self.work = torch.empty(max(lib.lsk_potrf_work_doubles(self.n), 1), dtype=F64, device=a.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=a.device)
        if self.n:
            _lib.check(lib.lsk_potrf_upper(_lib.ptr(a), self.n, a.stride(0), _lib.ptr(self.work), _lib.ptr(self._info),
                                           _lib.stream_ptr()), "lsk_potrf_upper")
            _lib.count_launch(3 * ((self.n + 127) // 128) - 2)
        if check and self.info:
            raise NotPositiveDefiniteError("matrix is not positive definite: pivot %d is not positive" % self.info)

    @property
    def info(self) -> int:
        """0, or the 1-based index of the first non-positive pivot (LAPACK dpotrf convention); synchronises."""
        return int(self._info.item())

    @property
    def U(self) -> torch.Tensor:
        return torch.triu(self.matrix)

    @property
    def L(self) -> torch.Tensor:
        return torch.triu(self.matrix).mT

    def logdet(self) -> torch.Tensor:
        return 2.0 * torch.log(torch.diagonal(self.matrix)).sum()

    def _solve(self, b: torch.Tensor, which: int) -> torch.Tensor:
        lib = _lib.load()
        if b.shape[0] != self.n or b.dim() > 2:
            raise ValueError("right-hand side must be [n] or [n, k] with n = %d" % self.n)
        vec = b.dim() == 1
        cols = b.to(F64).reshape(self.n, -1).t().contiguous().clone()      # [k, n]: one contiguous vector per column
        tmp = torch.empty(self.n + 128, dtype=F64, device=self.matrix.device)
        for c in range(cols.shape[0] if self.n else 0):
            _lib.check(lib.lsk_potrs_upper(_lib.ptr(self.matrix), self.n, self.matrix.stride(0),
                                           _lib.ptr(self.work), _lib.ptr(cols[c]), _lib.ptr(tmp), which, _lib.stream_ptr()),
                       "lsk_potrs_upper")
            _lib.count_launch(2 if which == 3 else 1)
        out = cols.t().contiguous()
        return out[:, 0] if vec else out

    def many_rhs(self) -> int:
        """From this many right-hand sides on, `solve` goes through the dense inverse: the inverse costs 2 n^3 / 3 flop once (99 ms at
        n = 16384, 2.9 ms at 4096) against one pipelined pair of sweeps per column (1.25 ms / 0.42 ms), so the break-even grows
        with n; a cached inverse is always used."""
        if getattr(self, "_inverse_panel", None) is not None:
            return 2
        return max(8, self.n // 256)

    def solve(self, b: torch.Tensor) -> torch.Tensor:
        """A^{-1} b (LAPACK dpotrs).  A vector or a few columns: two pipelined sweeps per column (`lsk_potrs_upper`: one launch per
        sweep, 1.25 ms at n = 16384).  Many columns: the backward sweep U X = Y has no blocked many-right-hand-side kernel, so the
        dense inverse (`lsk_potri_upper`, 2 n^3 / 3 flop on the tensor-core tile kernel, cached on the factor) is applied with one
        Gram launch, X = A^{-1} B."""
        if b.dim() == 2 and b.shape[1] >= self.many_rhs() and self.n:
            from . import ops
            if b.shape[0] != self.n:
                raise ValueError("right-hand side must have %d rows" % self.n)
            if getattr(self, "_inverse_panel", None) is None:
                self._inverse_panel = ops.PanelMatrix.from_dense(self.inverse())
            return ops.gram(self._inverse_panel, ops.PanelMatrix.from_dense(b.to(F64).t().contiguous()))
        return self._solve(b, 3)

    def solve_lower(self, b: torch.Tensor) -> torch.Tensor:
        """L^{-1} b = U^{-T} b (the whitening solve of the posterior variance).  A matrix of right-hand sides goes through
        the blocked `lsk_trsm_lower` (tensor-core tile kernel); a vector through the per-block matrix-vector steps."""
        if b.dim() == 2 and b.shape[1] > 4 and self.n:
            lib = _lib.load()
            if b.shape[0] != self.n:
                raise ValueError("right-hand side must have %d rows" % self.n)
            out = b.to(F64).contiguous().clone()
            work = torch.empty(lib.lsk_trsm_work_doubles(self.n, out.shape[1]), dtype=F64, device=out.device)
            _lib.check(lib.lsk_trsm_lower(_lib.ptr(self.matrix), self.n, self.matrix.stride(0), _lib.ptr(self.work), _lib.ptr(out),
                                          out.shape[1], out.stride(0), _lib.ptr(work), _lib.stream_ptr()), "lsk_trsm_lower")
            _lib.count_launch(3 * ((self.n + 127) // 128))
            return out
        return self._solve(b, 1)

    def solve_upper(self, b: torch.Tensor) -> torch.Tensor:
        """U^{-1} b."""
        return self._solve(b, 2)

    def inverse(self) -> torch.Tensor:
        """A^{-1} as a dense symmetric matrix (needed by the gradient of the marginal likelihood, tr(A^{-1} dA)):
        `lsk_potri_upper`, 2 n^3 / 3 flop on the tensor-core tile kernel."""
        lib = _lib.load()
        out = torch.empty((self.n, self.n), dtype=F64, device=self.matrix.device)
        if self.n:
            work = torch.empty(lib.lsk_potri_work_doubles(self.n), dtype=F64, device=self.matrix.device)
            _lib.check(lib.lsk_potri_upper(_lib.ptr(self.matrix), self.n, self.matrix.stride(0), _lib.ptr(self.work),
                                           _lib.ptr(work), _lib.ptr(out), self.n, _lib.stream_ptr()), "lsk_potri_upper")
            _lib.count_launch(4 * ((self.n + 127) // 128) + 1)
        return out


def cholesky(a: torch.Tensor, upper: bool = False) -> torch.Tensor:
    """Drop-in for `torch.linalg.cholesky(a, upper=upper)` on one float64 CUDA matrix."""
    f = CholeskyFactor(a)
    return f.U if upper else f.L.contiguous()
