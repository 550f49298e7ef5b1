"""SO(n) — drop-in for the reference's `spaces/so.py` on the kernel/sampling path.

Irreps, dimensions, Casimir values and their ordering come from `characters.py` (restating so.py:102-122,
171-207); characters are evaluated by the fused CUDA kernels through `classfn.GroupPlan`, never by
eigen-decomposition.  `rand` reproduces the reference's sampling algorithm (so.py:62-100) on the GPU."""
from __future__ import annotations

import math

import numpy as np
import torch

from ..space import CompactLieGroup, cuda_device

dtype = torch.float64


class SO(CompactLieGroup):
    """SO(n), special orthogonal group of degree `n` (so.py:19-41)."""

    family = "SO"

    def __init__(self, n: int, order=20):
        if n <= 2 and order:
            raise ValueError("Dimensions 1, 2 are not supported")
        self.n = n
        self.dim = n * (n - 1) // 2
        self.rank = n // 2
        self.order = order
        self.rho = np.arange(self.rank - 1, -1, -1) if n % 2 == 0 else np.arange(self.rank - 1, -1, -1) + 0.5
        self.id = torch.eye(n, device=cuda_device(), dtype=dtype).view(1, n, n)
        if order:
            CompactLieGroup.__init__(self, order=order)
        else:
            self.plan = None
            self.lb_eigenspaces = []

    def difference(self, x, y):
        return x @ y.T

    def rand(self, num=1):
        """Haar samples [num, n, n] (so.py:62-100).  Same sequence of torch RNG calls as the reference on the same
        device, so seeded scripts reproduce; QR + sign fixes run in `lsk_haar_qr` when available."""
        dev = cuda_device()
        n = self.n
        if n == 2:
            th = 2 * math.pi * torch.rand((num, 1), dtype=dtype, device=dev)
            c, s = torch.cos(th), torch.sin(th)
            return torch.cat((torch.hstack((c, s)).unsqueeze(-2), torch.hstack((-s, c)).unsqueeze(-2)), dim=-2)
        if n == 3:
            q = torch.randn((num, 4), dtype=dtype, device=dev)
            return quaternion_to_so3(q)
        h = torch.randn((num, n, n), device=dev, dtype=dtype)
        return haar_from_gaussian(h)

    @staticmethod
    def inv(x: torch.Tensor):
        return torch.transpose(x, -2, -1)

    def torus_representative(self, x):
        """Torus representative with the reference's conventions (so.py:136-168).  API-compatibility path: the
        fused kernels never call it."""
        n = self.n
        if n == 3:
            re = (torch.einsum('...ii->...', x) - 1) / 2
            im = torch.sqrt(torch.clamp(1 - torch.square(re), min=0.0))
            return torch.complex(re, im).unsqueeze(-1)
        if n % 2 == 1:
            ev = torch.linalg.eigvals(x)
            order = torch.sort(torch.view_as_real(ev), dim=-2).indices[..., 0]
            ev = torch.gather(ev, dim=-1, index=order)
            return ev[..., 0:-1:2]
        # even n: each spectrum belongs to two conjugacy classes; the reference separates them with the orientation of
        # the eigenvector frame (so.py:151-168).  Same construction here (API-compatibility path only).
        ev, vec = torch.linalg.eig(x)
        order = torch.sort(torch.view_as_real(ev), dim=-2).indices[..., 0]
        ev = torch.gather(ev, dim=-1, index=order)
        vec = torch.gather(vec, dim=-1, index=order.unsqueeze(-2).broadcast_to(vec.shape))
        frame = torch.zeros_like(vec)
        frame[..., ::2] = vec[..., ::2] + vec[..., 1::2]
        frame[..., 1::2] = vec[..., ::2] - vec[..., 1::2]
        frame = (frame.real + frame.imag) / math.sqrt(2)
        first = torch.pow(ev[..., 0], torch.det(frame).sgn())
        return torch.cat((first.unsqueeze(-1), ev[..., 2::2]), dim=-1)

    def pairwise_dist(self, x, y):
        emb = self.pairwise_embed(x, y)
        ang = torch.arccos(emb.real)
        d = math.sqrt(2) * torch.minimum(ang, 2 * math.pi - ang)
        return torch.norm(d, dim=1).reshape(x.shape[0], y.shape[0])


def quaternion_to_so3(q):
    """Unit-quaternion parametrisation used by the reference for SO(3) (so.py:74-92); `q` is the [num, 4] Gaussian draw.
    CUDA kernel `lsk_quaternion_to_group`."""
    from .. import _lib
    lib = _lib.load()
    q = q.to(dtype).contiguous()
    out = torch.empty((q.shape[0], 3, 3), dtype=dtype, device=q.device)
    _lib.check(lib.lsk_quaternion_to_group(_lib.ptr(q), q.shape[0], 0, _lib.ptr(out), _lib.stream_ptr()), "lsk_quaternion_to_group")
    _lib.count_launch()
    return out


def haar_from_gaussian(h):
    """QR of a Gaussian batch, column signs from diag(R), determinant fixed on column 0 (so.py:94-100):
    CUDA kernel `lsk_haar_from_gaussian` (one thread per matrix, LAPACK Householder conventions)."""
    from .. import _lib
    lib = _lib.load()
    h = h.to(dtype).contiguous()
    out = torch.empty_like(h)
    _lib.check(lib.lsk_haar_from_gaussian(_lib.ptr(h), h.shape[0], h.shape[-1], 0, _lib.ptr(out), _lib.stream_ptr()),
               "lsk_haar_from_gaussian")
    _lib.count_launch()
    return out
