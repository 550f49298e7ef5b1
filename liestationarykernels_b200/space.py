"""Manifold base classes — drop-in for the reference's `space.py` on the kernel/sampling path.

Same class names, attributes and call signatures (`rand`, `pairwise_diff`, `pairwise_embed`,
`torus_representative`, `inv`, `lb_eigenspaces[i].{index,dimension,lb_eigenvalue,phase_function}`), but the
O(N*M) work is never done here: kernels and samplers go through `classfn.GroupPlan` + the CUDA library, which
needs neither the pairwise matrix products nor the eigenvalues.  The methods below that *return* those
intermediates exist for API compatibility (the reference's tests and notebooks call them directly) and are
plain batched torch ops on the GPU.
"""
from __future__ import annotations

import math
import warnings
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib
from .classfn import group_plan

dtype = torch.float64


def cuda_device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class AbstractManifold(ABC):
    """Compact Lie group, compact homogeneous space or non-compact symmetric space (space.py:13-24)."""

    def __init__(self):
        ABC.__init__(self)

    @abstractmethod
    def rand(self, n=1):
        raise NotImplementedError

    def pairwise_diff(self, x, y):
        raise NotImplementedError


# ----------------------------------------------------------------------------------------------------------
# eigenspaces and characters of compact Lie groups
# ----------------------------------------------------------------------------------------------------------
class LieGroupCharacter(torch.nn.Module):
    """chi_lambda as a module (space.py:243-262).  `chi(gammas)` takes torus representatives like the reference;
    `evaluate(x)` takes group elements; `forward(gammas)` = dimension * chi(gammas)."""

    def __init__(self, *, representation):
        super().__init__()
        self.representation = representation

    @property
    def coeffs(self):
        cs, _ = self.representation.manifold._reference_table(self.representation.index)
        return cs

    @property
    def monoms(self):
        _, ms = self.representation.manifold._reference_table(self.representation.index)
        return ms

    def chi(self, gammas):
        return self.representation.manifold._chi_from_gammas(self.representation._position, gammas)

    def evaluate(self, x):
        return self.representation.manifold._chi_from_elements(self.representation._position, x)

    def forward(self, gammas):
        return self.representation.dimension * self.chi(gammas)


class LBEigenspace:
    """One irreducible representation = one Laplace-Beltrami eigenspace (space.py:163-219, so.py:171-207)."""

    def __init__(self, index, dimension, lb_eigenvalue, position, manifold):
        self.index = index
        self.dimension = dimension
        self.lb_eigenvalue = lb_eigenvalue
        self.manifold = manifold
        self._position = position
        self._phase_function = None

    def compute_dimension(self):
        return self.dimension

    def compute_lb_eigenvalue(self):
        return self.lb_eigenvalue

    def compute_phase_function(self):
        return LieGroupCharacter(representation=self)

    @property
    def phase_function(self):
        if self._phase_function is None:
            self._phase_function = self.compute_phase_function()
        return self._phase_function

    @property
    def basis(self):
        """Orthonormal basis of the eigenspace from translated characters (space.py:204-217, so.py:206-207, su.py:135-136);
        built on first use (dim^2 translations, a dim^2 x dim^2 Gram and its Cholesky factor: for small irreps only)."""
        if getattr(self, "_basis", None) is None:
            self._basis = self.compute_basis()
        return self._basis

    def compute_basis(self):
        return TranslatedCharactersBasis(representation=self)


class CompactLieGroup(AbstractManifold, ABC):
    """space.py:27-85.  Subclasses set `family` ('SO'/'SU'), `n`, `order` before calling this constructor."""

    family = None

    def __init__(self, *, order: int):
        super().__init__()
        self.plan = group_plan(self.family, self.n, order)
        self.lb_eigenspaces = [LBEigenspace(sig, dim, lam, pos, self) for pos, (sig, dim, lam) in enumerate(self.plan.irreps)]
        if order:
            error_est = sum(math.exp(-lam) * dim ** 2 for (_, dim, lam) in self.plan.dropped)
            if error_est > 1e-3:
                warnings.warn('Heat kernel error est. {}, consider larger order of approximation.'.format(error_est), stacklevel=3)
            self._lambdas = torch.tensor([r[2] for r in self.plan.irreps], dtype=dtype, device=cuda_device())
            self._dims = np.array([r[1] for r in self.plan.irreps], dtype=np.float64)

    @staticmethod
    @abstractmethod
    def inv(x):
        raise NotImplementedError

    @abstractmethod
    def torus_representative(self, x):
        raise NotImplementedError

    # -- API-compatibility paths (not used by the fused kernels) -------------------------------------------
    def pairwise_diff(self, x, y, reverse=False):
        """x_i y_j^{-1} (or x_i^{-1} y_j) for all pairs, flattened [N*M, n, n] with index i*M + j (space.py:62-78)."""
        if not reverse:
            out = torch.matmul(x[:, None], self.inv(y)[None, :])
        else:
            out = torch.matmul(self.inv(x)[:, None], y[None, :])
        return out.reshape(-1, self.n, self.n)

    def pairwise_embed(self, x, y):
        return self.torus_representative(self.pairwise_diff(x, y))

    def _phase_features(self, x, phases, row_scales, panel):
        """Phi[n, l*P + p] = row_scales[l] * d_l * Re chi_l(phase_p x_n^{-1})  (space.py:80-85 + 259-262 fused)."""
        from . import ops
        scales = np.asarray(row_scales, dtype=np.float64) * self._dims
        return ops.class_features(self.plan, x, phases, scales, panel)

    def _reference_table(self, signature):
        from . import characters as ch
        cs, ms = ch.reference_table_entry(self.family, self.n, tuple(signature))
        dev = cuda_device()
        return torch.tensor(cs, dtype=torch.int, device=dev), torch.tensor(ms, dtype=torch.int, device=dev)

    def _chi_from_gammas(self, position, gammas):
        """chi of one irrep at torus elements `gammas` [..., r] (SO) / [..., n] (SU): the weight-monomial sum of the
        reference (so.py:288-293, su.py:175-179), vectorised over monomials.  API-compatibility path only."""
        cs, ms = self._reference_table(self.plan.irreps[position][0])
        g = gammas.to(torch.complex128)
        if self.family == "SO":
            g = torch.cat((g, g.conj()), dim=-1)
        lead = g.shape[:-1]
        g = g.reshape(-1, 1, g.shape[-1])
        val = torch.zeros(g.shape[0], dtype=torch.complex128, device=g.device)
        step = 256
        for lo in range(0, ms.shape[0], step):
            terms = torch.prod(g ** ms[lo:lo + step][None], dim=-1)
            val += terms @ cs[lo:lo + step].to(torch.complex128)
        return val.reshape(lead)

    def _chi_from_elements(self, position, x):
        return self._chi_from_gammas(position, self.torus_representative(x))


class TranslatedCharactersBasis(torch.nn.Module):
    """Orthonormal basis of the space of matrix coefficients of one irrep V: the dim(V)^2 functions x -> d chi(g_m x) for random
    translations g_m, orthogonalised with the inverse Cholesky factor of their Gram matrix G[m, m'] = d chi(g_m g_m'^{-1})
    (space.py:284-314; same RNG call, `manifold.rand(dim ** 2)`, same retry rule: up to three draws).

    Not a hot path (the reference's EigenbasisKernel is the only caller and works for small irreps only: the Gram is
    dim^2 x dim^2): batched torch ops on the GPU through the API-compatibility methods (`pairwise_embed`, `torus_representative`,
    the weight-monomial character)."""

    def __init__(self, *, representation):
        super().__init__()
        self.representation = representation
        dim = self.representation.dimension
        self.character = self.representation.phase_function
        manifold = self.representation.manifold
        attempts = 0
        while True:
            attempts += 1
            self.translations = manifold.rand(dim ** 2)
            try:
                self._factorise()
                break
            except RuntimeError:
                if attempts >= 3:
                    raise

    def _factorise(self):
        """coeffs = inverse Cholesky factor of the Gram of the current translations (space.py:300-303)"""
        manifold, dim = self.representation.manifold, self.representation.dimension
        ratios = manifold.pairwise_embed(self.translations, self.translations)
        gram = self.character.forward(ratios).reshape(dim ** 2, dim ** 2)
        self.coeffs = torch.linalg.inv(torch.linalg.cholesky(gram))

    def forward(self, x):
        """[N, dim^2] complex: b_k(x_n) = sum_m coeffs[k, m] d chi(g_m x_n)"""
        manifold = self.representation.manifold
        x_translates = torch.matmul(self.translations.unsqueeze(0), x.unsqueeze(1))          # [N, dim^2, n, n]
        characters = self.character.forward(manifold.torus_representative(x_translates))     # [N, dim^2]
        return torch.matmul(self.coeffs, characters.to(self.coeffs.dtype).T).T


# ----------------------------------------------------------------------------------------------------------
# non-compact symmetric spaces (base)
# ----------------------------------------------------------------------------------------------------------
class NonCompactSymmetricSpace(AbstractManifold, ABC):
    """space.py:317-346"""

    def __init__(self):
        super().__init__()

    def dist(self, x, y):
        raise NotImplementedError

    def generate_lb_eigenspaces(self, measure):
        raise NotImplementedError
