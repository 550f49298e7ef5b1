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
    sp = sp / torch.linalg.vector_norm(sp, dim=-1, keepdim=True)
    a = torch.view_as_complex(sp[..., :2].contiguous()).unsqueeze(-1)
    b = torch.view_as_complex(sp[..., 2:].contiguous()).unsqueeze(-1)
    r1 = torch.hstack((a, -b.conj())).unsqueeze(-1)
    r2 = torch.hstack((b, a.conj())).unsqueeze(-1)
    return torch.cat((r1, r2), -1)


def haar_from_gaussian_complex(h):
    q, r = torch.linalg.qr(h)
    rd = torch.diagonal(r, dim1=-2, dim2=-1)
    q = q * torch.conj(rd / torch.abs(rd))[:, None]
    qd = torch.det(q)
    q[:, :, 0] *= torch.conj(qd / torch.abs(qd))[:, None]
    return q
