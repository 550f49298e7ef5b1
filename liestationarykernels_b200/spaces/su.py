"""SU(n) — drop-in for the reference's `spaces/su.py` on the kernel/sampling path (see spaces/so.py for the design)."""
from __future__ import annotations

import numpy as np
import torch

from ..space import CompactLieGroup, cuda_device

dtype = torch.cdouble


class SU(CompactLieGroup):
    """SU(n), special unitary group of degree `n` (su.py:20-37)."""

    family = "SU"

    def __init__(self, n: int, order=10):
        self.n = n
        self.dim = n * n - 1
        self.rank = n - 1
        self.order = order
        self.rho = np.arange(n - 1, -n, -2) * 0.5
        self.id = torch.eye(n, device=cuda_device(), dtype=dtype).view(1, n, n)
        if order:
            CompactLieGroup.__init__(self, order=order)
        else:
            self.plan = None
            self.lb_eigenspaces = []

    def dist(self, x, y):
        raise NotImplementedError

    def difference(self, x, y):
        return x @ y.mH

    def rand(self, num=1):
        """Haar samples [num, n, n] complex128 (su.py:45-64)."""
        dev = cuda_device()
        n = self.n
        if n == 2:
            sp = torch.randn((num, 4), dtype=torch.double, device=dev)
            return quaternion_to_su2(sp)
        h = torch.randn((num, n, n), dtype=dtype, device=dev)
        return haar_from_gaussian_complex(h)

    @staticmethod
    def inv(x: torch.Tensor):
        return torch.conj(torch.transpose(x, -2, -1))

    def torus_representative(self, x):
        """su.py:91-92 (API-compatibility path; not used by the fused kernels)."""
        return torch.linalg.eigvals(x)


def quaternion_to_su2(sp):
    """su.py:47-55; CUDA kernel `lsk_quaternion_to_group`."""
    from .. import _lib
    lib = _lib.load()
    sp = sp.to(torch.float64).contiguous()
    out = torch.empty((sp.shape[0], 2, 2), dtype=dtype, device=sp.device)
    _lib.check(lib.lsk_quaternion_to_group(_lib.ptr(sp), sp.shape[0], 1, _lib.ptr(torch.view_as_real(out)), _lib.stream_ptr()),
               "lsk_quaternion_to_group")
    _lib.count_launch()
    return out


def haar_from_gaussian_complex(h):
    """su.py:56-64; CUDA kernel `lsk_haar_from_gaussian` (complex Householder QR, unit-phase fixes)."""
    from .. import _lib
    lib = _lib.load()
    h = h.to(dtype).contiguous()
    out = torch.empty_like(h)
    _lib.check(lib.lsk_haar_from_gaussian(_lib.ptr(torch.view_as_real(h)), h.shape[0], h.shape[-1], 1,
                                          _lib.ptr(torch.view_as_real(out)), _lib.stream_ptr()), "lsk_haar_from_gaussian")
    _lib.count_launch()
    return out
