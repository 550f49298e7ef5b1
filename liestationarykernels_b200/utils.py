"""Small host helpers with the reference's names (`utils.py`): only what the kernel/sampling path uses."""
from __future__ import annotations

import torch


def hook_content_formula(lmd, n):
    """Dimension of the GL(n) irrep of highest weight `lmd` (utils.py:99-109); used by Stiefel.compute_inv_dimension."""
    numer, denom = 1, 1
    cols = [sum(row >= i + 1 for row in lmd) for i in range(lmd[0])] if lmd and lmd[0] > 0 else []
    for ir, row in enumerate(lmd):
        for ic in range(row):
            numer *= n + ic - ir
            denom *= cols[ic] + row - ir - ic - 1
    return numer / denom


def cartesian_prod(x1: torch.Tensor, x2: torch.Tensor):
    """utils.py:40-52, as broadcast *views* (nothing is materialised): [a, b, ...] for both operands."""
    a, b = x1.shape[0], x2.shape[0]
    return x1[:, None].expand(a, b, *x1.shape[1:]), x2[None, :].expand(a, b, *x2.shape[1:])


def lazy_property(fn):
    attr = "_lazy_" + fn.__name__

    @property
    def _lazy(self):
        if not hasattr(self, attr):
            setattr(self, attr, fn(self))
        return getattr(self, attr)
    return _lazy


def GOE_sampler(num_samples, n):
    """Spectra of GOE matrices (utils.py:7-11): same torch RNG call as the reference; eigvalsh stays a library call
    (setup-time, m small matrices)."""
    from .space import cuda_device
    s = torch.randn(num_samples, n, n, device=cuda_device(), dtype=torch.float64)
    s = (s + torch.transpose(s, -2, -1)) / 2
    return torch.linalg.eigvalsh(s, UPLO='U')
