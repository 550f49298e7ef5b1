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
        return self.scale * (self.left.to_dense() * self.right.to_dense()).sum(dim=1)

    def __getitem__(self, idx):
        return self.evaluate()[idx]


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
