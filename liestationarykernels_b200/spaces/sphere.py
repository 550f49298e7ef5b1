"""S^n and the projective space RP^n with zonal phase functions — drop-in for `spaces/sphere.py:17-218`
(SURVEY.md 8f row f4).

The degree-l eigenspace contributes  phi_l(x, y) = d_l C_l^alpha(<x, y>) / C_l^alpha(1),  alpha = (n - 1) / 2  (Gegenbauer),
so the truncated kernel is ONE polynomial of the dot product.  The reference evaluates every degree separately on the
materialised N*M dot products, in the monomial basis (`GegenbauerPolynomials`, sphere.py:170-203; it cancels to ~1e-10 at
degree 20).  Here the weighted sum is re-expanded on the host in the Chebyshev basis (exact quadrature in extended
precision, once per space) and evaluated by Clenshaw's recurrence in the epilogue of the pairwise tile kernel, whose
tensor-core stage forms the dot products: `lsk_pairwise_poly` with one bilinear form of K = n + 1 and a CHEB1 / FEAT_CHEB1
epilogue — the same kernel that serves SO(3)."""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
import torch

from .. import _lib, ops
from ..classfn import GroupPlan
from ..space import AbstractManifold, LBEigenspace, cuda_device

dtype = torch.float64
REF_PI = 3.1415927410125732          # 2 * float32(acos(0)), the module constant of the reference (sphere.py:15)


def num_harmonics(dimension: int, degree: int) -> int:
    """Number of spherical harmonics of `degree` in R^dimension (what the reference imports from `spherical_harmonics`)."""
    if degree == 0:
        return 1
    if dimension == 3:
        return 2 * degree + 1
    return int(Fraction(2 * degree + dimension - 2, degree) * math.comb(degree + dimension - 3, degree - 1))


RECURRENCE_FROM = 25     # sphere.py:149-152: degrees >= 25 switch from monomial coefficients to `NewGegenbauerPolynomials`


def _gegenbauer_table(alpha: float, max_degree: int, t: np.ndarray, c0=1) -> np.ndarray:
    """C_m^alpha(t) for m = 0..max_degree by the three-term recurrence, extended precision; [max_degree + 1, len(t)].
    c0 = 0 reproduces `NewGegenbauerPolynomials.forward` (sphere.py:209-217), which seeds the recurrence with C_0 = 0
    instead of 1: for degrees >= 25 the reference's phase function is that (different) polynomial, normalised to d_l at
    t = 1 — kept for parity (quirk G10)."""
    t = t.astype(np.longdouble)
    a = np.longdouble(alpha)
    out = np.zeros((max_degree + 1, len(t)), dtype=np.longdouble)
    out[0] = c0
    if max_degree >= 1:
        out[1] = 2 * a * t
    for m in range(2, max_degree + 1):
        out[m] = (2 * t * (m + a - 1) * out[m - 1] - (m + 2 * a - 2) * out[m - 2]) / m
    return out


class ZonalSphericalFunction(torch.nn.Module):
    """sphere.py:146-167: phase function of one eigenspace on dot products, d_l C_l(t) / C_l(1) (API-compatibility path:
    plain torch ops on whatever device `dist` lives on)."""

    def __init__(self, dim, n, chebyshev_row):
        super().__init__()
        self.dim, self.n = dim, n
        self._row = chebyshev_row

    def forward(self, dist):
        c = torch.as_tensor(self._row, dtype=dtype, device=dist.device)
        b1 = torch.zeros_like(dist)
        b2 = torch.zeros_like(dist)
        for k in range(len(c) - 1, 0, -1):
            b1, b2 = c[k] + 2 * dist * b1 - b2, b1
        return c[0] + dist * b1 - b2


class SphereLBEigenspace(LBEigenspace):
    def compute_phase_function(self):
        m = self.manifold
        row = m.plan.M[self._position]
        return ZonalSphericalFunction(m.dim, self.index, row[: self.index + 1])

    @property
    def basis(self):
        """Orthonormal basis of the spherical harmonics of this degree (sphere.py:126-141), built on first use."""
        if getattr(self, "_basis", None) is None:
            self._basis = NormalizedSphericalFunctions(self.manifold.dim, self.index, self.phase_function, self.dimension)
        return self._basis


class NormalizedSphericalFunctions(torch.nn.Module):
    """An orthonormal basis of the degree-`degree` spherical harmonics on S^dim (sphere.py:133-141).

    The reference delegates to `spherical_harmonics.SphericalHarmonicsLevel` (vdutor/SphericalHarmonics, absent from the image),
    whose published construction is: take a fundamental system v_1..v_N of N = dim H_l points, form the Gram matrix of the zonal
    functions G_ij = Z_l(<v_i, v_j>), factor G = L L^T, and return Y(x) = L^{-1} [Z_l(<v_i, x>)]_i.  Restated here with
    Z_l = this package's phase function (N_l C_l(t) / C_l(1), so that sum_k Y_k(x) Y_k(y) = Z_l(<x, y>), the addition theorem).
    That package ships precomputed, optimised fundamental systems; here the system is chosen greedily (pivoted Cholesky over a
    fixed pseudo-random candidate set, its own generator: the global RNG is not touched).  A different fundamental system gives a
    different orthonormal basis of the SAME space, so basis values are not comparable entry by entry with the reference's, only
    basis-independent quantities are (the sum over the basis, i.e. everything a kernel uses)."""

    def __init__(self, dimension, degree, phase_function, num_harmonics):
        super().__init__()
        self.dimension, self.degree, self.zonal, self.num = dimension, degree, phase_function, int(num_harmonics)
        dev = cuda_device()
        gen = torch.Generator(device="cpu").manual_seed(1000003 * dimension + degree)
        cand = torch.randn(4 * self.num + 64, dimension + 1, generator=gen, dtype=dtype)
        cand = (cand / torch.linalg.norm(cand, dim=1, keepdim=True)).to(dev)
        gram = self.zonal(cand @ cand.T)
        # greedy pivoted Cholesky: pick the candidate with the largest remaining diagonal
        diag = torch.diagonal(gram).clone()
        rows = torch.zeros((self.num, cand.shape[0]), dtype=dtype, device=dev)
        chosen = []
        for k in range(self.num):
            j = int(torch.argmax(diag))
            chosen.append(j)
            r = (gram[j] - rows[:k, j] @ rows[:k]) / torch.sqrt(diag[j])
            rows[k] = r
            diag = diag - r * r
            diag[chosen] = -1.0
        self.fundamental_system = cand[chosen].contiguous()
        g = self.zonal(self.fundamental_system @ self.fundamental_system.T)
        self.chol_inv = torch.linalg.inv(torch.linalg.cholesky(g))

    def forward(self, x):
        """[N, num_harmonics]"""
        z = self.zonal(self.fundamental_system @ x.to(dtype).T)          # [num, N]
        return (self.chol_inv @ z).T


class _ZonalPlan(GroupPlan):
    """The slice of `classfn.GroupPlan` the rank-1 (Clenshaw) epilogues need: one bilinear form = the dot product."""

    def __init__(self, n: int, degrees, dims):
        self.family, self.n, self.order = "Sphere", n, len(degrees)
        self.kind = "cheb1"
        self.is_complex = False
        self.square_mask = 0
        self.nforms = 1
        self.kg_total = (n + 1 + 3) // 4
        self.group_begin, self.group_end = [0], [self.kg_total]
        self.aff = [[0.0, 1.0]]                               # the Chebyshev variable is the dot product itself
        self.irreps = [(d, dims[i], d * (n + d - 1)) for i, d in enumerate(degrees)]
        self.dropped = []
        deg = max(degrees) if degrees else 0
        q = deg + 1                                           # Chebyshev-Gauss nodes: exact for degree <= deg
        theta = (np.arange(q, dtype=np.longdouble) + np.longdouble(0.5)) * (np.arccos(np.longdouble(-1.0)) / q)
        nodes = np.cos(theta)
        alpha = (n - 1) / 2.0
        geg = _gegenbauer_table(alpha, deg, nodes)
        geg1 = _gegenbauer_table(alpha, deg, np.ones(1))[:, 0]
        if deg >= RECURRENCE_FROM:
            geg[RECURRENCE_FROM:] = _gegenbauer_table(alpha, deg, nodes, c0=0)[RECURRENCE_FROM:]
            geg1[RECURRENCE_FROM:] = _gegenbauer_table(alpha, deg, np.ones(1), c0=0)[RECURRENCE_FROM:, 0]
        cheb = np.cos(np.arange(q, dtype=np.longdouble)[:, None] * theta[None, :])      # T_k at the nodes
        M = np.zeros((len(degrees), q))
        for l, d in enumerate(degrees):
            vals = geg[d] / geg1[d] * np.longdouble(dims[l])
            coef = (cheb @ vals) * (np.longdouble(2.0) / q)
            coef[0] /= 2
            coef[d + 1:] = 0                                   # exact zeros above the degree
            coef[(d + 1) % 2::2] = 0                           # and for the wrong parity
            M[l] = coef.astype(np.float64)
        self.M = M


class Sphere(AbstractManifold):
    """S^n in R^{n+1} (sphere.py:17-57).  `order` = number of eigenspaces (degrees 0 .. order - 1)."""

    _degree_step = 1

    def __init__(self, n: int, order=10):
        super().__init__()
        if n < 2:
            raise ValueError("Sphere / ProjectiveSpace need n >= 2 (the Gegenbauer parameter (n - 1) / 2 must be positive)")
        self.n = self.dim = n
        self.order = order
        degrees = [self._degree_step * index for index in range(order)]
        if degrees and max(degrees) + 1 > _lib.MAX_COEF:
            raise ValueError("order too large: degree %d exceeds the epilogue's %d Chebyshev coefficients" % (max(degrees), _lib.MAX_COEF))
        dims = [num_harmonics(n + 1, d) for d in degrees]
        self.plan = _ZonalPlan(n, degrees, dims)
        self.lb_eigenspaces = [SphereLBEigenspace(d, dims[pos], d * (n + d - 1), pos, self) for pos, d in enumerate(degrees)]
        dev = cuda_device()
        self._lambdas = torch.tensor([float(d * (n + d - 1)) for d in degrees], dtype=dtype, device=dev)
        self._dims = np.ones(len(degrees))                    # d_l is already inside the phase functions (rows of plan.M)
        self.id = torch.zeros((1, n + 1), device=dev, dtype=dtype)
        self.id[0][0] = 1.0

    # ---- reference API -----------------------------------------------------------------------------------------
    def dist(self, x, y):
        return torch.arccos(torch.dot(x, y))

    def rand(self, n=1):
        if n == 0:
            return None
        x = torch.randn(n, self.n + 1, dtype=dtype, device=cuda_device())
        return x / torch.norm(x, dim=1, keepdim=True)

    def pairwise_embed(self, x, y):
        """all dot products, flattened [N * M] with index i * M + j (sphere.py:48-54)"""
        return (x[:, None, :] * y[None, :, :]).sum(dim=-1).reshape(-1)

    def pairwise_dist(self, x, y):
        return torch.abs(torch.arccos(self.pairwise_embed(x, y))).reshape((x.shape[0], y.shape[0]))

    # ---- fused path --------------------------------------------------------------------------------------------
    def _operand(self, x):
        _lib.require_cuda()
        if x.dim() != 2 or x.shape[1] != self.n + 1:
            raise ValueError("expected [num, %d] points, got %s" % (self.n + 1, tuple(x.shape)))
        if not x.is_cuda:
            raise _lib.LskError("inputs must live on the CUDA device (no CPU fallback)")
        return ops.PanelMatrix.from_dense(x.to(dtype))

    def _eigensum_coefs(self, x, y, coefs, scale, out=None):
        """scale * sum_l coefs[l] phi_l(<x_i, y_j>) with one launch of the pairwise kernel."""
        ep = self.plan.epilogue(np.asarray(coefs, dtype=np.float64), scale)
        same = y is x
        a = self._operand(x)
        b = a if same else self._operand(y)
        return ops.pairwise_poly(a.buf, b.buf, x.shape[0], y.shape[0], self.plan.kg_total, ep, out=out, symmetric=same)

    def _eigensum(self, kernel, x, y, normalize, out=None):
        from ..spectral_kernel import spectral_weights
        w = spectral_weights(kernel.measure, self)
        scale = float(torch.abs(kernel.measure.variance[0]).item() / float(kernel.normalizer)) if normalize else 1.0
        return self._eigensum_coefs(x, y, w, scale, out)

    def _identity_values(self):
        """phi_l at coinciding points = d_l"""
        return torch.tensor([float(e.dimension) for e in self.lb_eigenspaces], dtype=dtype, device=self._lambdas.device)

    def _phase_features(self, x, phases, row_scales, panel):
        """Phi[n, l * P + p] = row_scales[l] phi_l(<phase_p, x_n>)  (prior_approximation.py:70-83 on the sphere)."""
        a, b = self._operand(x), self._operand(phases)
        return ops.features_from_operands(self.plan, a.buf, b.buf, x.shape[0], phases.shape[0], row_scales, panel, x.device)


class ProjectiveSpace(Sphere):
    """RP^n = S^n / {+-1}: even degrees only (sphere.py:60-108)."""

    _degree_step = 2

    def dist(self, x, y):
        d = torch.arccos(torch.clip(torch.dot(x, y), -1, 1))
        return torch.min(d, REF_PI - d)

    def pairwise_dist(self, x, y):
        d = torch.arccos(torch.clip(self.pairwise_embed(x, y), -1.0, 1.0))
        return torch.min(d, REF_PI - d).reshape((x.shape[0], y.shape[0]))
