"""The flat torus T^n = [0,1)^n with glued boundary — drop-in for `spaces/torus.py:14-97`.

Irreps are integer frequency vectors k (dimension 1, eigenvalue 4 pi^2 |k|^2); the truncated character sum
    k(x, y) = sum_k a(lambda_k) cos(2 pi k.(x - y))
is evaluated as the Gram of a 2 * order-column feature map (CUDA kernel `lsk_torus_features`) on the fp64 tensor-core
Gram kernel: nothing of size N * M * order is formed (the reference materialises the N * M differences and loops over
the irreps)."""
from __future__ import annotations

import itertools
import math
import warnings

import numpy as np
import torch

from .. import _lib, ops
from ..space import AbstractManifold, LBEigenspace, cuda_device
from ..utils import cartesian_prod

dtype = torch.float64
REF_PI = 3.1415927410125732          # 2 * float32(acos(0)): the module constant the reference's characters use (torus.py:11)


class TorusCharacter(torch.nn.Module):
    """torus.py:88-95: chi_k(x) = exp(2 pi i k.x) on raw differences x (the torus is its own maximal torus)."""

    def __init__(self, *, representation):
        super().__init__()
        self.representation = representation

    def chi(self, x):
        k = torch.tensor(self.representation.index, dtype=dtype, device=x.device)
        return torch.pow(torch.e, 2 * REF_PI * 1j * torch.sum(x * k[:, ...], dim=1))

    def evaluate(self, x):
        return self.chi(x)

    def forward(self, x):
        return self.representation.dimension * self.chi(x)


class TorusLBEigenspace(LBEigenspace):
    def compute_phase_function(self):
        return TorusCharacter(representation=self)


class Torus(AbstractManifold):
    def __init__(self, n: int, order=20):
        super().__init__()
        self.n, self.dim, self.order = n, n, order
        self.id = torch.zeros(n, dtype=dtype, device=cuda_device()).view(1, n)
        self.lb_eigenspaces = []
        if order:
            spaces = [(sig, 4 * math.pi * math.pi * sum(v ** 2 for v in sig)) for sig in self.generate_signatures(order)]
            spaces.sort(key=lambda t: t[1])                  # space.py:36-38: stable sort on the eigenvalue, first `order`
            kept, dropped = spaces[:order], spaces[order:]
            self.lb_eigenspaces = [TorusLBEigenspace(sig, 1, lam, pos, self) for pos, (sig, lam) in enumerate(kept)]
            error_est = sum(math.exp(-lam) for _, lam in dropped)
            if error_est > 1e-3:
                warnings.warn('Heat kernel error est. {}, consider larger order of approximation.'.format(error_est), stacklevel=2)
            self._lambdas = torch.tensor([lam for _, lam in kept], dtype=dtype, device=cuda_device())
            self._dims = np.ones(len(kept))
            self._freq = torch.tensor([sig for sig, _ in kept], dtype=dtype, device=cuda_device()).contiguous()

    # ---- reference API -----------------------------------------------------------------------------------------
    def generate_signatures(self, order):
        vals = range(-order // 2, order // 2 + 1)
        return list(itertools.product(vals, repeat=self.n))

    def rand(self, n=1):
        return torch.rand((n, self.n), device=cuda_device(), dtype=dtype)

    def difference(self, x, y):
        return x - y

    @staticmethod
    def inv(x):
        return -x

    def torus_representative(self, x):
        return x

    def dist(self, x, y):
        diff = torch.abs(x - y)
        return torch.norm(torch.minimum(diff, 1 - diff), dim=1)

    def pairwise_embed(self, x, y):
        x_, y_ = cartesian_prod(x, y)
        return torch.reshape(x_, (-1, self.n)) - torch.reshape(y_, (-1, self.n))

    pairwise_diff = pairwise_embed

    def pairwise_dist(self, x, y):
        diff = self.pairwise_embed(x, y)
        diff = torch.minimum(diff, 1 - diff)
        return torch.norm(diff, dim=1).reshape(x.shape[0], y.shape[0])

    # ---- fused path --------------------------------------------------------------------------------------------
    def _features(self, x, weights):
        """[N, 2 * order] panel-major feature map; `weights` (device tensor or None) multiplies the columns of irrep k."""
        _lib.require_cuda()
        lib = _lib.load()
        if x.dim() != 2 or x.shape[1] != self.n:
            raise ValueError("expected [num, %d] points, got %s" % (self.n, tuple(x.shape)))
        if not x.is_cuda:
            raise _lib.LskError("inputs must live on the CUDA device (no CPU fallback)")
        xc = x.to(dtype).contiguous()
        L = len(self.lb_eigenspaces)
        pm = ops.PanelMatrix(xc.shape[0], 2 * L, xc.device, buf=torch.empty(0, dtype=dtype, device=xc.device))
        pm.buf = torch.empty(max(pm.row_panels * pm.kg * 256, 1), dtype=dtype, device=xc.device)
        if xc.shape[0]:
            w = None if weights is None else weights.to(dtype).contiguous()
            _lib.check(lib.lsk_torus_features(_lib.ptr(xc), xc.shape[0], self.n, _lib.ptr(self._freq),
                                              None if w is None else _lib.ptr(w), L, 2 * REF_PI, _lib.ptr(pm.buf), pm.kg,
                                              _lib.stream_ptr()), "lsk_torus_features")
            _lib.count_launch()
        return pm

    def _eigensum_coefs(self, x, y, coefs, scale, out=None):
        """scale * sum_k coefs[k] cos(2 pi k.(x_i - y_j)) as one Gram."""
        w = torch.as_tensor(np.asarray(coefs, dtype=np.float64), device=x.device)
        fx = self._features(x, w)
        fy = self._features(y, None)
        return ops.gram(fx, fy, scale, out=out)

    def _eigensum(self, kernel, x, y, normalize, out=None):
        from ..spectral_kernel import spectral_weights
        w = spectral_weights(kernel.measure, self)
        scale = float(torch.abs(kernel.measure.variance[0]).item() / float(kernel.normalizer)) if normalize else 1.0
        return self._eigensum_coefs(x, y, w, scale, out)

    def _identity_values(self):
        return torch.ones(len(self.lb_eigenspaces), dtype=dtype, device=self._lambdas.device)
