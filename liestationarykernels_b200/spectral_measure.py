"""Spectral measures a(lambda) of the heat and Matern kernels — drop-in for the reference's
`spectral_measure.py:8-36` (same constructor arguments, same `nn.Parameter` attributes, same formulas).

The measures stay thin torch modules: per kernel call they are evaluated once on the vector of Casimir
eigenvalues (L numbers); everything of size N*M happens in the CUDA kernels."""
from __future__ import annotations

import torch
from torch.nn import Parameter

from . import _lib

dtype = torch.float64


def _device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class AbstractSpectralMeasure(torch.nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, eigenvalues):
        raise NotImplementedError


class MaternSpectralMeasure(AbstractSpectralMeasure):
    """a(lambda) = (2 nu / l^2 + lambda)^(-nu - dim/2)   (spectral_measure.py:17-26)"""

    def __init__(self, dim, lengthscale, nu, variance=1.0):
        super().__init__(dim)
        dev = _device()
        self.lengthscale = Parameter(torch.tensor([lengthscale], device=dev, dtype=dtype, requires_grad=True))
        self.variance = Parameter(torch.tensor([variance], device=dev, dtype=dtype, requires_grad=True))
        self.nu = torch.tensor([nu], device=dev, dtype=dtype)

    def forward(self, eigenvalue):
        ls = self.lengthscale
        return torch.pow(self.nu / (ls.clone() * ls.clone() / 2.0) + eigenvalue, -self.nu - self.dim / 2)


class SqExpSpectralMeasure(AbstractSpectralMeasure):
    """a(lambda) = exp(-l^2 lambda / 2)   (spectral_measure.py:29-36)"""

    def __init__(self, dim, lengthscale, variance=1.0):
        super().__init__(dim)
        dev = _device()
        self.lengthscale = Parameter(torch.tensor([lengthscale], device=dev, dtype=dtype, requires_grad=True))
        self.variance = Parameter(torch.tensor([variance], device=dev, dtype=dtype, requires_grad=True))

    def forward(self, eigenvalues):
        ls = self.lengthscale
        return torch.exp(-ls.clone() * ls.clone() / 2.0 * eigenvalues)


# north_star spellings
HeatSpectralMeasure = SqExpSpectralMeasure
