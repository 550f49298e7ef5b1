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
