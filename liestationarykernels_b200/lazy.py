"""Minimal stand-in for `gpytorch.lazy.MatmulLazyTensor(NonLazyTensor(Fx), NonLazyTensor(Fy).t())`, which is what the
reference's RandomPhaseKernel / RandomFourierFeatureKernel return (spectral_kernel.py:125-127, 195-197).
gpytorch is not in this image; the object keeps the two feature maps (in the tensor-core operand layout) and
evaluates `Fx Fy^T` on demand with the fp64 DMMA Gram kernel."""
from __future__ import annotations

import torch

from . import ops


class LazyGram:
    def __init__(self, left: ops.PanelMatrix, right: ops.PanelMatrix, scale: float = 1.0):
        self.left, self.right, self.scale = left, right, scale

    @property
    def shape(self):
        return torch.Size((self.left.rows, self.right.rows))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def evaluate(self, out=None):
        return ops.gram(self.left, self.right, self.scale, out=out)

    to_dense = evaluate

    def t(self):
        return LazyGram(self.right, self.left, self.scale)

    def clone(self):
        return self

    def matmul(self, rhs):
        """(Fx Fy^T) rhs without forming the Gram: Fx (Fy^T rhs)."""
        rhs2 = rhs if rhs.dim() == 2 else rhs[:, None]
        inner = ops.gram(ops.PanelMatrix.from_dense(rhs2.t().contiguous()), _transpose(self.right))   # [c, K]
        res = ops.gram(self.left, ops.PanelMatrix.from_dense(inner), self.scale)
        return res if rhs.dim() == 2 else res[:, 0]

    __matmul__ = matmul

    def diag(self):
        """diag(Fx Fy^T) = row-wise dot products, taken on the panel buffers (same layout on both sides, padding is zero)."""
        if self.left.rows != self.right.rows:
            raise ValueError("diag() needs a square Gram")
        lv = self.left.buf[: self.left.row_panels * self.left.kg * 256].view(self.left.row_panels, self.left.kg, 64, 4)
        rv = self.right.buf[: self.right.row_panels * self.right.kg * 256].view(self.right.row_panels, self.right.kg, 64, 4)
        return self.scale * (lv * rv).sum(dim=(1, 3)).reshape(-1)[: self.left.rows]

    def __getitem__(self, idx):
        return self.evaluate()[idx]


def _gpytorch_classes():
    """(matmul class, dense wrapper class) of an importable gpytorch: the 1.7-era `gpytorch.lazy` API the reference targets
    (spectral_kernel.py:9) or its successor `linear_operator`; None when neither is installed (this image)."""
    try:
        from gpytorch.lazy import MatmulLazyTensor, NonLazyTensor
        return MatmulLazyTensor, NonLazyTensor
    except Exception:
        pass
    try:
        from linear_operator.operators import DenseLinearOperator, MatmulLinearOperator
        return MatmulLinearOperator, DenseLinearOperator
    except Exception:
        return None


USE_GPYTORCH = True       # hand real gpytorch objects to callers when gpytorch is importable (set False to always get LazyGram)


def as_reference_lazy(gram: "LazyGram"):
    """What the reference returns, `MatmulLazyTensor(NonLazyTensor(Fx), NonLazyTensor(Fy).t())` (spectral_kernel.py:125-127,
    195-197), as a REAL gpytorch / linear_operator object when that package is importable, so that gpytorch models consume
    the kernel unchanged; the feature maps are handed over dense and row-major, which is gpytorch's layout (its matmul is then
    cuBLAS, not the DMMA Gram kernel - use `LazyGram.evaluate()` / `USE_GPYTORCH = False` for that).  Without gpytorch: `gram`."""
    classes = _gpytorch_classes() if USE_GPYTORCH else None
    if classes is None or gram.left is None:
        return gram
    matmul_cls, dense_cls = classes
    left = gram.left.to_dense()
    if gram.scale != 1.0:
        left = left * gram.scale
    right = left if (gram.right is gram.left and gram.scale == 1.0) else gram.right.to_dense()
    return matmul_cls(dense_cls(left), dense_cls(right).t())


def _transpose(pm: ops.PanelMatrix) -> ops.PanelMatrix:
    return ops.PanelMatrix.from_dense(pm.to_dense().t().contiguous())


class _PhaseGramFn(torch.autograd.Function):
    """K = sum_l Phi_l(x) Phi_l(y)^T with Phi_l = sqrt(c_l) Psi_l, as a differentiable function of the coefficient vector c
    (RandomPhaseKernel: c_l = |sigma^2| a_l(lengthscale) / (P norm)).  dL/dc_l = <G, Phi_l(x) Phi_l(y)^T> / c_l: one Gram of
    a column block per irrep in backward, nothing of size N * M * L is stored."""

    @staticmethod
    def forward(ctx, c, left, right, n_irreps, n_phases):
        ctx.left, ctx.right, ctx.n_irreps, ctx.n_phases = left, right, n_irreps, n_phases
        ctx.save_for_backward(c)
        return ops.gram(left, right)

    @staticmethod
    def backward(ctx, grad):
        (c,) = ctx.saved_tensors
        left, right, L, P = ctx.left, ctx.right, ctx.n_irreps, ctx.n_phases
        g = torch.zeros(L, dtype=torch.float64, device=grad.device)
        grad = grad.contiguous()
        if P % 4 == 0:
            for l in range(L):
                g[l] = (grad * ops.gram_columns(left, right, l * P // 4, (l + 1) * P // 4)).sum()
        else:               # column blocks do not start on a k-group: go through dense slices
            dl, dr = left.to_dense(), right.to_dense()
            for l in range(L):
                kl = ops.gram(ops.PanelMatrix.from_dense(dl[:, l * P:(l + 1) * P].contiguous()),
                              ops.PanelMatrix.from_dense(dr[:, l * P:(l + 1) * P].contiguous()))
                g[l] = (grad * kl).sum()
        safe = torch.where(c != 0, c, torch.ones_like(c))
        return torch.where(c != 0, g / safe, torch.zeros_like(g)), None, None, None, None


class DifferentiableLazyGram(LazyGram):
    """LazyGram whose `evaluate()` stays attached to the autograd graph of the kernel's hyper-parameters."""

    def __init__(self, left, right, coefficients, n_irreps, n_phases):
        super().__init__(left, right, 1.0)
        self.coefficients, self.n_irreps, self.n_phases = coefficients, n_irreps, n_phases

    def evaluate(self, out=None):
        k = _PhaseGramFn.apply(self.coefficients, self.left, self.right, self.n_irreps, self.n_phases)
        if out is not None:
            out.copy_(k.detach())
        return k

    to_dense = evaluate

    def t(self):
        return DifferentiableLazyGram(self.right, self.left, self.coefficients, self.n_irreps, self.n_phases)


class _FourierGramFn(torch.autograd.Function):
    """K = amp Re(F_x F_y^H) on a non-compact symmetric space as a differentiable function of the spectral samples lmd
    ([m] on H^n, [m, n] on SPD(n)), the c-function values coeff [m] and the amplitude amp = |sigma^2| / norm.
        dF_ik / dlmd_kr = i D_ikr F_ik,   D = -log((1 - |x|^2) / |x - b|^2) on H^n (hyperbolic.py:118-129),
                                          D_r = log a_r(U_i h_k^{-1}) on SPD(n) (spd.py:92-106)
    With the upstream gradient G and T = G conj(F_y), S = G^T F_x (four real Grams on the tile kernel):
        A_k  = Re sum_i F_ik T_ik                              -> d/d amp = sum_k A_k,  d/d coeff_k = 2 amp A_k / coeff_k
        B_kr = Im sum_i D_ikr F_ik T_ik,  C_kr = Im sum_j D_jkr conj(F_jk) S_jk      -> d/d lmd_kr = -amp (B_kr - C_kr)
    The lengthscale reaches lmd (scaled samples) and coeff (c-function) through ordinary torch autograd on the host."""

    @staticmethod
    def forward(ctx, lmd, coeff, amp, manifold, funcs, x, y, grouped=False):
        """grouped: x and y are already in group form (`manifold.to_group` applied: the pairwise differences of RandomSpectralKernel)"""
        ctx.manifold, ctx.funcs, ctx.x, ctx.y, ctx.grouped = manifold, funcs, x, y, grouped
        ctx.save_for_backward(lmd, coeff, amp)
        s = float(amp.detach()) ** 0.5
        group = (lambda t: t) if grouped else manifold.to_group
        px = manifold._features(group(x), funcs, s, "panel")
        py = px if y is x else manifold._features(group(y), funcs, s, "panel")
        return ops.gram(px, py)

    @staticmethod
    def backward(ctx, grad):
        lmd, coeff, amp = ctx.saved_tensors
        m, funcs, x, y = ctx.manifold, ctx.funcs, ctx.x, ctx.y
        group = (lambda t: t) if ctx.grouped else m.to_group
        fx = m._features(group(x), funcs, 1.0, "complex")
        fy = fx if y is x else m._features(group(y), funcs, 1.0, "complex")
        dx = m._log_terms(x, funcs, ctx.grouped)            # [N, m, R]
        dy = dx if y is x else m._log_terms(y, funcs, ctx.grouped)
        g = grad.contiguous()
        pg, pgt = ops.PanelMatrix.from_dense(g), ops.PanelMatrix.from_dense(g.t().contiguous())

        def times(pm, f):       # (dense operand in panel form) times the real / imaginary parts of f
            re = ops.gram(pm, ops.PanelMatrix.from_dense(f.real.t().contiguous()))
            im = ops.gram(pm, ops.PanelMatrix.from_dense(f.imag.t().contiguous()))
            return re, im

        t_re, t_im = times(pg, fy)                      # T = G conj(F_y) = t_re - i t_im
        s_re, s_im = times(pgt, fx)                     # S = G^T F_x    = s_re + i s_im
        ft = fx * torch.complex(t_re, -t_im)
        fs = fy.conj() * torch.complex(s_re, s_im)
        a_k = ft.real.sum(dim=0)
        b_kr = (dx * ft.imag[:, :, None]).sum(dim=0)
        c_kr = (dy * fs.imag[:, :, None]).sum(dim=0)
        g_amp = a_k.sum().reshape(amp.shape)
        g_coeff = 2.0 * amp * a_k / coeff
        g_lmd = (-amp * (b_kr - c_kr)).reshape(lmd.shape)
        return g_lmd, g_coeff, g_amp, None, None, None, None, None


class DifferentiableFourierGram(LazyGram):
    """Lazy Gram of the Fourier-feature kernel on H^n / SPD(n) whose `evaluate()` carries hyper-parameter gradients."""

    def __init__(self, manifold, funcs, x, y, lmd, coeff, amp, grouped=False):
        self.manifold, self.funcs, self.x, self.y = manifold, funcs, x, y
        self.lmd, self.coeff, self.amp, self.grouped = lmd, coeff, amp, grouped
        self.left = self.right = None
        self.scale = 1.0

    @property
    def shape(self):
        return torch.Size((self.x.shape[0], self.y.shape[0]))

    def evaluate(self, out=None):
        k = _FourierGramFn.apply(self.lmd, self.coeff, self.amp, self.manifold, self.funcs, self.x, self.y, self.grouped)
        if out is not None:
            out.copy_(k.detach())
        return k

    to_dense = evaluate

    def t(self):
        return DifferentiableFourierGram(self.manifold, self.funcs, self.y, self.x, self.lmd, self.coeff, self.amp, self.grouped)

    def __getitem__(self, idx):
        return self.evaluate()[idx]

    # the feature maps are built inside the autograd Function; the matrix-free members go through the dense result
    def matmul(self, rhs):
        return self.evaluate() @ rhs

    __matmul__ = matmul

    def diag(self):
        return torch.diagonal(self.evaluate())
