"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference hot path, torch float64/complex128 on CPU.

This file is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`liestationarykernels_b200/` imports it, and the product raises when its CUDA library is missing.

It follows the reference algorithm step by step — materialised cartesian products, `bmm`, LAPACK `eigvals`,
the per-monomial character loop — so that (a) it is a faithful timing stand-in for the reference's CPU path on
the GPU box (where `/root/reference` does not exist) and (b) its numbers agree with the real reference to
rounding.  Parity pinning: `oracle/make_golden.py` runs the *real* reference (via `oracle/ref_shim.py`) in the
build container and stores inputs+outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this port
against those files, and `tests/test_oracle_vs_reference.py` checks it live against the reference when the
tree is mounted.  Every function cites the reference lines it restates (paths relative to
`/root/reference/src/lie_stationary_kernels/`).

RNG: functions that draw random numbers issue the same torch calls in the same order as the reference, so a
seeded CPU generator produces bit-identical draws.
"""
from __future__ import annotations

import math
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

# Weight multiplicities (the content of the reference's precomputed_characters.json) and the irrep ordering
# come from the table generator, which tests/test_characters.py pins against a digest of the reference tables.
from liestationarykernels_b200 import characters as _ch  # noqa: E402

F64 = torch.float64
C128 = torch.complex128
# hyperbolic.py:11, spd.py:12: the reference builds pi from a float32 acos, i.e. 3.1415927410125732 (G9)
REF_PI = 2 * torch.acos(torch.zeros(1)).item()


# ----------------------------------------------------------------------------------------------------------
# utils.py
# ----------------------------------------------------------------------------------------------------------
def pair_grid(a: torch.Tensor, b: torch.Tensor):
    """utils.py:40-52 `cartesian_prod`: both operands materialised at [len(a), len(b), ...]."""
    reps_a = [1, b.shape[0]] + [1] * (a.dim() - 1)
    reps_b = [a.shape[0], 1] + [1] * (b.dim() - 1)
    return torch.tile(a[:, None, ...], reps_a), torch.tile(b[None, :, ...], reps_b)


def goe_spectrum(num: int, n: int):
    """utils.py:7-11 `GOE_sampler`."""
    s = torch.randn(num, n, n, dtype=F64)
    s = (s + s.transpose(-2, -1)) / 2
    return torch.linalg.eigvalsh(s, UPLO="U")


def hook_content(lmd, n):
    """utils.py:99-109 `hook_content_formula`."""
    numer, denom = 1, 1
    cols = [sum(row >= i + 1 for row in lmd) for i in range(lmd[0])] if lmd and lmd[0] > 0 else []
    for ir, row in enumerate(lmd):
        for ic in range(row):
            numer *= n + ic - ir
            denom *= cols[ic] + row - ir - ic - 1
    return numer / denom


# ----------------------------------------------------------------------------------------------------------
# spectral_measure.py
# ----------------------------------------------------------------------------------------------------------
class Measure:
    """kind 'matern' (spectral_measure.py:17-26) or 'heat' (:29-36); plain floats held as 1-element tensors."""

    def __init__(self, kind, dim, lengthscale, nu=None, variance=1.0):
        self.kind, self.dim = kind, dim
        self.lengthscale = torch.tensor([lengthscale], dtype=F64)
        self.variance = torch.tensor([variance], dtype=F64)
        self.nu = None if nu is None else torch.tensor([nu], dtype=F64)

    def __call__(self, eigenvalue):
        ls = self.lengthscale
        if self.kind == "matern":
            return torch.pow(self.nu / (ls * ls / 2.0) + eigenvalue, -self.nu - self.dim / 2)
        return torch.exp(-ls * ls / 2.0 * eigenvalue)


# ----------------------------------------------------------------------------------------------------------
# spaces/so.py, spaces/su.py — sampling, inverse, torus representative, characters
# ----------------------------------------------------------------------------------------------------------
class Group:
    """SO(n) / SU(n) with the first `order` irreps (space.py:29-44)."""

    def __init__(self, family: str, n: int, order: int):
        self.family, self.n, self.order = family, n, order
        if family == "SO":
            self.dim, self.rank = n * (n - 1) // 2, n // 2
            self.dtype = F64
        else:
            self.dim, self.rank = n * n - 1, n - 1
            self.dtype = C128
        self.id = torch.eye(n, dtype=self.dtype).view(1, n, n)
        if order:
            kept, _ = _ch.truncated_irreps(family, n, order)
        else:
            kept = []
        self.irreps = kept                          # (signature, dimension, casimir)
        self._tables = {}

    # so.py:62-100, su.py:45-64
    def rand(self, num=1):
        n = self.n
        if self.family == "SO":
            if n == 2:
                th = 2 * math.pi * torch.rand((num, 1), dtype=F64)
                c, s = torch.cos(th), torch.sin(th)
                return torch.cat((torch.hstack((c, s)).unsqueeze(-2), torch.hstack((-s, c)).unsqueeze(-2)), dim=-2)
            if n == 3:
                q = torch.randn((num, 4), dtype=F64)
                q /= torch.linalg.vector_norm(q, dim=-1, keepdim=True)
                x, y, z, w = (q[..., i].unsqueeze(-1) for i in range(4))
                xx, yy, zz = x ** 2, y ** 2, z ** 2
                xy, xz, xw, yz, yw, zw = x * y, x * z, x * w, y * z, y * w, z * w
                r1 = torch.hstack((1 - 2 * (yy + zz), 2 * (xy - zw), 2 * (xz + yw))).unsqueeze(-1)
                r2 = torch.hstack((2 * (xy + zw), 1 - 2 * (xx + zz), 2 * (yz - xw))).unsqueeze(-1)
                r3 = torch.hstack((2 * (xz - yw), 2 * (yz + xw), 1 - 2 * (xx + yy))).unsqueeze(-1)
                return torch.cat((r1, r2, r3), -1)
            h = torch.randn((num, n, n), dtype=F64)
            q, r = torch.linalg.qr(h)
            q *= torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))[:, None]
            q[:, :, 0] *= torch.sign(torch.det(q))[:, None]
            return q
        if n == 2:
            sp = torch.randn((num, 4), dtype=F64)
            sp /= torch.linalg.vector_norm(sp, dim=-1, keepdim=True)
            a = torch.view_as_complex(sp[..., :2]).unsqueeze(-1)
            b = torch.view_as_complex(sp[..., 2:]).unsqueeze(-1)
            r1 = torch.hstack((a, -b.conj())).unsqueeze(-1)
            r2 = torch.hstack((b, a.conj())).unsqueeze(-1)
            return torch.cat((r1, r2), -1)
        h = torch.randn((num, n, n), dtype=C128)
        q, r = torch.linalg.qr(h)
        rd = torch.diagonal(r, dim1=-2, dim2=-1)
        q *= torch.conj(rd / torch.abs(rd))[:, None]
        qd = torch.det(q)
        q[:, :, 0] *= torch.conj(qd / torch.abs(qd))[:, None]
        return q

    # so.py:124-127, su.py:79-82
    def inv(self, x):
        xt = torch.transpose(x, -2, -1)
        return xt if self.family == "SO" else torch.conj(xt)

    # space.py:62-78
    def pairwise_diff(self, x, y, reverse=False):
        if not reverse:
            a, b = pair_grid(x, self.inv(y))
        else:
            a, b = pair_grid(self.inv(x), y)
        n = self.n
        return torch.bmm(a.reshape(-1, n, n), b.reshape(-1, n, n))

    # so.py:136-168, su.py:91-92
    def torus_representative(self, x):
        if self.family == "SU":
            return torch.linalg.eigvals(x)
        n = self.n
        if n == 3:
            re = (torch.einsum("...ii->...", x) - 1) / 2
            im = torch.sqrt(torch.max(1 - torch.square(re), torch.zeros_like(re)))
            return torch.view_as_complex(torch.cat((re.unsqueeze(-1), im.unsqueeze(-1)), -1)).unsqueeze(-1)
        if n % 2 == 1:
            ev = torch.linalg.eigvals(x)
            order = torch.sort(torch.view_as_real(ev), dim=-2).indices[..., 0]
            ev = torch.gather(ev, dim=-1, index=order)
            return ev[..., 0:-1:2]
        ev, vec = torch.linalg.eig(x)
        order = torch.sort(torch.view_as_real(ev), dim=-2).indices[..., 0]
        ev = torch.gather(ev, dim=-1, index=order)
        vec = torch.gather(vec, dim=-1, index=order.unsqueeze(-2).broadcast_to(vec.shape))
        c = torch.zeros_like(vec)
        c[..., ::2] = vec[..., ::2] + vec[..., 1::2]
        c[..., 1::2] = vec[..., ::2] - vec[..., 1::2]
        c = c.real + c.imag
        c /= math.sqrt(2)
        torch.pow(ev[..., 0], torch.det(c).sgn(), out=ev[..., 0])
        return ev[..., ::2]

    # space.py:80-85
    def pairwise_embed(self, x, y):
        return self.torus_representative(self.pairwise_diff(x, y))

    def _table(self, signature):
        """(coeffs, exponents) int tensors in the JSON layout (so.py:214-223 / su.py:143-153)."""
        key = tuple(signature)
        if key not in self._tables:
            cs, ms = _ch.reference_table_entry(self.family, self.n, key)
            self._tables[key] = (torch.tensor(cs, dtype=torch.int), torch.tensor(ms, dtype=torch.int))
        return self._tables[key]

    # so.py:288-293, su.py:175-179
    def chi(self, signature, gammas):
        cs, ms = self._table(signature)
        if self.family == "SO":
            gammas = torch.cat((gammas, gammas.conj()), dim=-1)
        val = torch.zeros(gammas.shape[:-1], dtype=C128)
        for c, m in zip(cs, ms):
            val += c * torch.prod(gammas ** m, dim=-1)
        return val

    # space.py:259-262  (LieGroupCharacter.forward)
    def phase_function(self, irrep, gammas):
        sig, dim, _ = irrep
        return dim * self.chi(sig, gammas)


# ----------------------------------------------------------------------------------------------------------
# spaces/stiefel.py, spaces/grassmannian.py, space.py:88-143, :265-281
# ----------------------------------------------------------------------------------------------------------
def _block_diag(hu, hd, nu, nd):
    z = torch.zeros((hu.shape[0], nu, nd), dtype=F64)
    left = torch.cat((hu, z.transpose(-1, -2)), dim=-2)
    right = torch.cat((z, hd), dim=-2)
    return torch.cat((left, right), dim=-1)


class Homogeneous:
    """kind in {'stiefel', 'grassmannian', 'oriented_grassmannian'}: SO(n)/H with H sampled `average_order` times."""

    def __init__(self, kind, n, m, order=10, average_order=None):
        assert n > m
        self.kind, self.n, self.m, self.n_m = kind, n, m, n - m
        if average_order is None:
            average_order = 50 if kind == "oriented_grassmannian" else 30
        self.average_order = average_order
        self.g = Group("SO", n, order)
        self.order = order
        hdim = {"stiefel": self.n_m * (self.n_m - 1) // 2,
                "oriented_grassmannian": m * (m - 1) // 2 + self.n_m * (self.n_m - 1) // 2,
                "grassmannian": m * (m - 1) // 2 + self.n_m * (self.n_m - 1) // 2 - 1}[kind]
        self.dim = self.g.dim - hdim
        self.h_samples = self.sample_h(average_order)          # space.py:97
        self.irreps = self.g.irreps
        self.id = torch.zeros((n, m), dtype=F64).fill_diagonal_(1.0).view(1, n, m)

    def sample_h(self, num):
        n, m, nm = self.n, self.m, self.n_m
        if self.kind == "stiefel":                              # stiefel.py:25-36 (H = SO(n-m), order 0)
            raw = Group("SO", nm, 0).rand(num)
            eye = torch.eye(m, dtype=F64).view(1, m, m).repeat(num, 1, 1)
            return _block_diag(eye, raw, m, nm)
        if self.kind == "oriented_grassmannian":                # grassmannian.py:12-27
            hu = Group("SO", m, 0).rand(num)
            hd = Group("SO", nm, 0).rand(num)
            return _block_diag(hu, hd, m, nm)
        assert num % 2 == 0                                     # grassmannian.py:29-55
        hu = Group("SO", m, 0).rand(num // 2)
        hd = Group("SO", nm, 0).rand(num // 2)
        first = _block_diag(hu, hd, m, nm).clone()
        hu[:, :, -1] = hu[:, :, -1] * -1
        hd[:, :, -1] = hd[:, :, -1] * -1
        return torch.cat((first, _block_diag(hu, hd, m, nm)), dim=0)

    def rand(self, num=1):                                      # space.py:121-123 + G_to_M
        return self.g.rand(num)[:, :, :self.m]

    def lift(self, x):                                          # stiefel.py:38-48 `M_to_G`
        g, r = torch.linalg.qr(x, mode="complete")
        rd = torch.diagonal(r, dim1=-2, dim2=-1)
        rd = torch.cat((rd, torch.ones((x.shape[0], self.n_m), dtype=F64)), dim=1)
        g = g * rd[:, None]
        diff = 2 * (torch.all(torch.isclose(g[:, :, :self.m], x), dim=-1).type(F64) - 0.5)
        g = g * diff[..., None]
        g[:, :, -1] *= torch.sign(torch.det(g))[:, None]
        assert torch.allclose(x, g[:, :, :x.shape[-1]])
        return g

    def pairwise_diff(self, x, y):                              # space.py:125-129
        return self.g.pairwise_diff(self.lift(x), self.lift(y), reverse=True)

    def pairwise_embed(self, x, y):                             # space.py:131-135
        return self.g.pairwise_embed(self.pairwise_diff(x, y), self.h_samples)

    def phase_function(self, irrep, gammas_xh):
        """AveragedLieGroupCharacter.forward (space.py:273-275): d * chi averaged over the H samples."""
        val = self.g.phase_function(irrep, gammas_xh).reshape(-1, self.average_order)
        return torch.mean(val, dim=-1)

    def compute_inv_dimension(self, signature):                 # stiefel.py:63-69 / grassmannian.py
        if self.kind != "stiefel":
            return 1
        m_ = min(self.m, self.n_m)
        if m_ < self.g.rank and signature[m_] > 0:
            return 0
        return hook_content(tuple(abs(v) for v in signature), m_)


# ----------------------------------------------------------------------------------------------------------
# spaces/torus.py — the flat torus [0,1)^n as a compact Lie group
# ----------------------------------------------------------------------------------------------------------
class Torus:
    """torus.py:14-97.  Irreps are integer frequency vectors k; chi_k(x) = exp(2 pi i k.x), dimension 1, eigenvalue
    4 pi^2 |k|^2 (exact pi); the character itself uses the reference's module constant `pi = 2 * acos(0)` evaluated in
    float32 (torus.py:11), i.e. 3.1415927410125732."""

    def __init__(self, n, order=20):
        import itertools
        self.n, self.dim, self.order = n, n, order
        self.id = torch.zeros(n, dtype=F64).view(1, n)
        vals = range(-order // 2, order // 2 + 1)                           # torus.py:55-57
        sigs = list(itertools.product(vals, repeat=n))
        spaces = [(sig, 1, 4 * math.pi * math.pi * sum(v ** 2 for v in sig)) for sig in sigs]    # torus.py:78-82
        spaces.sort(key=lambda t: t[2])                                     # space.py:36-38: stable sort, first `order`
        self.irreps = spaces[:order] if order else []

    def rand(self, num=1):                                                  # torus.py:52-53
        return torch.rand((num, self.n), dtype=F64)

    def pairwise_embed(self, x, y):                                         # torus.py:39-43
        xa, ya = pair_grid(x, y)
        return xa.reshape(-1, self.n) - ya.reshape(-1, self.n)

    def phase_function(self, irrep, diff):                                  # torus.py:91-95 + space.py:260-262 (d = 1)
        k = torch.tensor(irrep[0], dtype=F64)
        t = torch.sum(diff * k[:, ...], dim=1)
        return torch.pow(torch.e, 2 * REF_PI * 1j * t)

    def dist(self, x, y):                                                   # torus.py:32-36
        d = torch.abs(x - y)
        return torch.norm(torch.minimum(d, 1 - d), dim=1)

    def pairwise_dist(self, x, y):                                          # torus.py:45-50 (no abs: as the reference)
        d = self.pairwise_embed(x, y)
        d = torch.minimum(d, 1 - d)
        return torch.norm(d, dim=1).reshape(x.shape[0], y.shape[0])


# ----------------------------------------------------------------------------------------------------------
# spaces/sphere.py — S^n and RP^n with zonal (Gegenbauer) phase functions
# ----------------------------------------------------------------------------------------------------------
class Sphere:
    """sphere.py:17-57 (Sphere) and :60-108 (ProjectiveSpace: even degrees only).  irreps = (degree, dimension, eigenvalue).
    The dimension comes from the absent `spherical_harmonics.num_harmonics` (published formula restated in
    oracle/ref_shim.py); the phase function uses the reference's own constant exp(log d_n) / C_n(1) (sphere.py:159-163)."""

    def __init__(self, n, order=10, projective=False):
        self.n, self.dim, self.order, self.projective = n, n, order, projective
        self.id = torch.zeros((1, n + 1), dtype=F64)
        self.id[0][0] = 1.0
        self.irreps = []
        for index in range(order):
            deg = 2 * index if projective else index
            self.irreps.append((deg, self._num_harmonics(n + 1, deg), deg * (n + deg - 1)))       # sphere.py:122-127

    @staticmethod
    def _num_harmonics(dimension, degree):
        if degree == 0:
            return 1
        if dimension == 3:
            return int(2 * degree + 1)
        return int(round((2 * degree + dimension - 2) / degree * math.comb(degree + dimension - 3, degree - 1)))

    def rand(self, num=1):                                                  # sphere.py:41-46
        x = torch.randn(num, self.n + 1, dtype=F64)
        return x / torch.norm(x, dim=1, keepdim=True)

    def pairwise_embed(self, x, y):                                         # sphere.py:48-54: all dot products, index i*M + j
        xa, ya = pair_grid(x, y)
        return (xa.reshape(-1, self.n + 1) * ya.reshape(-1, self.n + 1)).sum(dim=1)

    def pairwise_dist(self, x, y):                                          # sphere.py:56-57 / :104-107
        if self.projective:
            d = torch.arccos(torch.clip(self.pairwise_embed(x, y), -1.0, 1.0))
            return torch.min(d, REF_PI - d).reshape(x.shape[0], y.shape[0])
        return torch.abs(torch.arccos(self.pairwise_embed(x, y))).reshape(x.shape[0], y.shape[0])

    def _gegenbauer(self, deg, t):
        """sphere.py:170-218: monomial coefficients through loggamma for degree < 25, three-term recurrence above."""
        from scipy.special import loggamma
        import numpy as np
        alpha = (self.dim - 1) / 2.0
        if deg < 25:
            coef = torch.zeros(deg + 1, dtype=F64)
            if deg == 0:
                coef[0] = 1
            if deg == 1:
                coef[1] = 2 * alpha
            if deg >= 2:
                for k in range(0, deg // 2 + 1):
                    log_c = (deg - 2 * k) * np.log(2) + loggamma(deg - k + alpha) - loggamma(alpha) - loggamma(k + 1) \
                        - loggamma(deg - 2 * k + 1)
                    coef[deg - 2 * k] = (-1) ** k * np.exp(log_c)
            powers = torch.arange(0., deg + 1., dtype=F64)
            return (torch.pow(t[..., None], powers) * coef).sum(dim=-1)
        c = [0 * t, 2 * alpha * t]                                          # sphere.py:209-217 (starts from C_0 = 0 as written)
        for m in range(2, deg + 1):
            c.append((c[-1] * 2 * t * (m + alpha - 1) - (m + 2 * alpha - 2) * c[-2]) / m)
        return c[-1]

    def phase_function(self, irrep, emb):                                   # sphere.py:146-167
        from scipy.special import loggamma
        import numpy as np
        deg = irrep[0]
        if deg == 0:
            const = 1.0
        else:
            log_d = np.log(2 * deg + self.dim - 1) + loggamma(deg + self.dim - 1) - loggamma(self.dim) - loggamma(deg + 1)
            const = float(np.exp(log_d)) / float(self._gegenbauer(deg, torch.tensor(1.0, dtype=F64)))
        return self._gegenbauer(deg, emb) * torch.tensor([const], dtype=F64)[0]


# ----------------------------------------------------------------------------------------------------------
# spectral_kernel.py:26-54 — EigenbasisSumKernel
# ----------------------------------------------------------------------------------------------------------
def eigensum_unnormalised(space, measure, x, y):
    """spectral_kernel.py:45-54 with normalize=False: sum_l a(l) * Re phase_function_l(embed(x, y))."""
    emb = space.pairwise_embed(x, y)
    cov = torch.zeros(len(x), len(y), dtype=F64)
    for irrep in space.irreps:
        cov += measure(irrep[2]) * space.phase_function(irrep, emb).view(x.size()[0], y.size()[0]).real
    return cov.real


def eigensum_normalizer(space, measure):
    """spectral_kernel.py:27-35: the un-normalised kernel at one random point (consumes RNG like the reference)."""
    p = space.rand()
    return eigensum_unnormalised(space, measure, p, p)[0, 0]


def eigensum(space, measure, x, y, normalizer):
    return torch.abs(measure.variance[0]) * eigensum_unnormalised(space, measure, x, y) / normalizer


# ----------------------------------------------------------------------------------------------------------
# spectral_kernel.py:78-127 — RandomPhaseKernel;  prior_approximation.py:50-92 — RandomPhaseApproximation
# ----------------------------------------------------------------------------------------------------------
def random_phase_embedding(space, measure, phases, x, scale_fn):
    """spectral_kernel.py:96-109 / prior_approximation.py:70-83: [N, L*P], irrep-major columns.
    `scale_fn(a_lambda)` is sqrt(a) for the kernel and sqrt(|var| a)/sqrt(norm) for the sampler."""
    P = phases.shape[0]
    emb = space.pairwise_embed(phases, x)
    cols = []
    for irrep in space.irreps:
        e = space.phase_function(irrep, emb).real.view(P, x.size()[0]).T
        e = scale_fn(measure(irrep[2])) * e
        cols.append(e / math.sqrt(P))
    return torch.cat(cols, dim=1)


def random_phase_kernel_from_embeddings(measure, ex, ey, normalizer=None):
    """spectral_kernel.py:119-127 on ready feature maps (sqrt(a)-scaled, `random_phase_embedding(..., torch.sqrt)`)."""
    if normalizer is not None:
        s = torch.sqrt(torch.abs(measure.variance[0]))
        ex = s * ex / torch.sqrt(normalizer)
        ey = s * ey / torch.sqrt(normalizer)
    return ex @ ey.t().clone()


def random_phase_kernel(space, measure, phases, x, y, normalizer=None):
    """spectral_kernel.py:111-127; returns the dense Gram the lazy tensor would evaluate to."""
    ex = random_phase_embedding(space, measure, phases, x, torch.sqrt)
    ey = random_phase_embedding(space, measure, phases, y, torch.sqrt)
    return random_phase_kernel_from_embeddings(measure, ex, ey, normalizer)


def random_phase_sample(space, measure, phases, weights, x, normalizer):
    """prior_approximation.py:70-88."""
    var = torch.abs(measure.variance[0])

    def scale(a):
        return torch.sqrt(var * a)

    emb = random_phase_embedding(space, measure, phases, x, scale) / torch.sqrt(normalizer)
    return torch.einsum("nm,m->n", emb, weights)


def random_phase_sample_from_embedding(measure, emb_sqrt_a, weights, normalizer):
    """prior_approximation.py:85-88 on a ready sqrt(a)-scaled feature map: sqrt(|var| a) = sqrt(|var|) sqrt(a)."""
    emb = torch.sqrt(torch.abs(measure.variance[0])) * emb_sqrt_a / torch.sqrt(normalizer)
    return torch.einsum("nm,m->n", emb, weights)


# ----------------------------------------------------------------------------------------------------------
# spaces/hyperbolic.py
# ----------------------------------------------------------------------------------------------------------
class Hyperbolic:
    def __init__(self, n, order=10000):
        self.n, self.dim, self.order = n, n, order
        self.id = torch.zeros(n, dtype=F64).view(1, n)
        self.normalized_lmd = None
        self.shift = None

    def rand_phase(self, num=1):                                # hyperbolic.py:71-77
        x = torch.randn(num, self.n, dtype=F64)
        return x / torch.norm(x, dim=1, keepdim=True)

    def rand(self, num=1):                                      # hyperbolic.py:79-84
        sphere = self.rand_phase(num)
        r = torch.pow(torch.rand(num, dtype=F64), 1 / self.n) * 0.99
        return sphere * r[:, None].clone()

    def to_group(self, x):                                      # hyperbolic.py:54-55
        return x.squeeze(dim=-1)

    def generate(self, measure):
        """hyperbolic.py:27-52: draws lambda (once, cached) and the shifts; returns (lmd, shift)."""
        if self.normalized_lmd is None:
            if measure.kind == "matern":
                df = 2 * measure.nu + measure.dim
                st = torch.distributions.StudentT(df=df - 1, scale=1).rsample((self.order,))
                lmd = torch.abs(st) / torch.sqrt(df - 1)
            else:
                nrm = torch.distributions.Normal(torch.tensor(0, dtype=F64), torch.tensor(1, dtype=F64))
                lmd = torch.abs(nrm.rsample((self.order,)).type(F64))
            self.normalized_lmd = torch.squeeze(lmd)
            self.shift = self.rand_phase(self.order)
        if measure.kind == "matern":
            scale = torch.sqrt(1 / 4 + 2 * measure.nu[0] / (measure.lengthscale[0] ** 2))
        else:
            scale = 1.0 / torch.abs(measure.lengthscale[0])
        return self.normalized_lmd * scale, self.shift

    def c_function(self, lmd):                                  # hyperbolic.py:134-150
        n = self.n
        if n % 2 == 0:
            adds = torch.tensor([(2 * i + 1) ** 2 / 4 for i in range(n // 2 - 1)], dtype=F64)
        else:
            adds = torch.tensor([i ** 2 for i in range(n // 2)], dtype=F64)
        log_c = torch.sum(torch.log(torch.square(lmd)[:, None].clone() + adds[None, :].clone()), dim=1)
        if n % 2 == 0:
            log_c = log_c + torch.log(lmd) + torch.log(torch.tanh(REF_PI * lmd))
        return torch.squeeze(torch.exp(log_c / 2))

    def features(self, x, lmd, shift):
        """hyperbolic.py:118-129 and :153-158: [N, m] complex128."""
        rho = torch.tensor([(self.n - 1) / 2], dtype=F64)[0]
        xa, sa = pair_grid(x, shift)
        d = torch.sum(torch.square(xa.reshape(-1, self.n) - sa.reshape(-1, self.n)), dim=1)
        den = torch.log(d).reshape(x.size()[0], -1)
        num = torch.log(1 - torch.sum(torch.square(x), dim=1))
        log_xb = num[:, None].clone() - den
        inner = torch.einsum("nm,m-> nm", log_xb, -1j * lmd + rho)
        return torch.einsum("nm,m->nm", torch.exp(inner), self.c_function(lmd).type(C128))

    def pairwise_diff(self, x, y):                              # hyperbolic.py:57-69
        xa, ya = pair_grid(x, y)
        xf, yf = xa.reshape(-1, self.n), ya.reshape(-1, self.n)
        d2 = torch.sum(torch.square(xf - yf), dim=1)
        nx, ny = torch.sum(torch.square(xf), dim=1), torch.sum(torch.square(yf), dim=1)
        dist = torch.arccosh(1 + 2 * d2 / (1 - nx) / (1 - ny))
        ones = torch.ones((dist.size()[0], self.n), dtype=F64) / math.sqrt(self.n)
        return ones * torch.tanh(dist / 2)[:, None].clone()


# ----------------------------------------------------------------------------------------------------------
# spaces/spd.py, space.py:338-380
# ----------------------------------------------------------------------------------------------------------
class SPD:
    def __init__(self, n, order=200):
        self.n, self.dim, self.order = n, n * (n + 1) // 2, order
        self.id = torch.eye(n, dtype=F64).view(1, n, n)

    def rand_phase(self, num=1):                                # spd.py:48-55
        g = torch.randn((num, self.n, self.n), dtype=F64)
        q, r = torch.linalg.qr(g)
        q *= torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))[:, None]
        q[:, :, 0] *= torch.sign(torch.det(q))[:, None]
        return q

    def rand(self, num=1):                                      # spd.py:57-64
        g = torch.randn(num, self.n, self.n, dtype=F64)
        return torch.bmm(g, g.transpose(-2, -1)) * 4 / (math.factorial(self.n + 1) ** (1 / self.n))

    def to_group(self, x):                                      # spd.py:45-46
        return torch.linalg.cholesky(x, upper=True)

    def generate(self, measure):
        """spd.py:27-43: re-samples shifts and spectra on every call; returns (lmd [m, n], shift [m, n, n])."""
        shift = self.rand_phase(self.order)
        if measure.kind == "matern":
            nu, ls = measure.nu[0], measure.lengthscale[0]
            scale = 1 / torch.sqrt((self.n ** 3 - self.n) / 48 + 2 * nu / ls)
            goe = goe_spectrum(self.order, self.n)
            chi = torch.sqrt(torch.distributions.chi2.Chi2(2 * nu).rsample((self.order,)))
            lmd = goe / chi[:, None] / scale
        else:
            lmd = goe_spectrum(self.order, self.n) / measure.lengthscale
        return lmd, shift

    def c_function(self, lmd):                                  # spd.py:116-123
        n = self.n
        iu = torch.triu_indices(n, n, offset=1)
        diff = (lmd[:, None, :] - lmd[:, :, None])[:, iu[0], iu[1]]
        return torch.exp(torch.sum(torch.log(torch.tanh(REF_PI * torch.abs(diff))), dim=1) / 2)

    def features(self, u, lmd, shift):
        """space.py:370-380 with spd.py:92-106: u = to_group(x) [N, n, n]; -> [N, m] complex128."""
        n = self.n
        a, b = pair_grid(u, torch.linalg.inv(shift))            # space.py:338-346
        prod = torch.bmm(a.reshape(-1, n, n), b.reshape(-1, n, n))
        _, an = torch.linalg.qr(prod, mode="complete")          # spd.py:93
        sign = torch.diag_embed(torch.diagonal(torch.sign(an), dim1=-2, dim2=-1))
        an = torch.bmm(sign, an)
        diag = torch.diagonal(an, dim1=-2, dim2=-1)
        log_a = torch.log(diag).type(C128).view(u.shape[0], lmd.shape[0], -1)
        rho = torch.tensor([(i + 1) - (n + 1) / 2 for i in range(n)], dtype=F64)
        lin = 1j * lmd + rho[None, :]
        expo = torch.exp(torch.einsum("nmr,mr->nm", log_a, lin))
        return torch.einsum("nm,m->nm", expo, self.c_function(lmd).type(C128))


# ----------------------------------------------------------------------------------------------------------
# spectral_kernel.py:164-197 — RandomFourierFeatureKernel;  prior_approximation.py:95-123
# ----------------------------------------------------------------------------------------------------------
def rff_normalizer(space, lmd, shift):
    """spectral_kernel.py:168-173: the un-normalised kernel at the identity point, entry [0,0]."""
    f = space.features(space.to_group(space.id), lmd, shift)
    psi = torch.cat((f.real.clone(), f.imag.clone()), dim=-1)
    return (psi @ psi.t().clone())[0, 0]


def rff_kernel(space, measure, lmd, shift, x, y, normalizer=None):
    """spectral_kernel.py:175-197: dense Gram of [Re F, Im F]."""
    fx = space.features(space.to_group(x), lmd, shift)
    fy = space.features(space.to_group(y), lmd, shift)
    if normalizer is not None:
        s = torch.sqrt(torch.abs(measure.variance[0]))
        fx = s * fx / torch.sqrt(normalizer)
        fy = s * fy / torch.sqrt(normalizer)
    px = torch.cat((fx.real.clone(), fx.imag.clone()), dim=-1)
    py = torch.cat((fy.real.clone(), fy.imag.clone()), dim=-1)
    return px @ py.t().clone()


def random_spectral_kernel(space, measure, lmd, shift, x, y, normalizer=None):
    """spectral_kernel.py:140-161 (RandomSpectralKernel.forward): Re F(x_i y_j^{-1}) conj(F(id))^T; H^n only (`pairwise_diff`)."""
    diff = space.pairwise_diff(space.to_group(x), space.to_group(y))
    f_diff = space.features(diff, lmd, shift)
    f_id = space.features(space.to_group(space.id), lmd, shift)
    cov = (f_diff @ torch.conj(f_id).T).view(x.size()[0], y.size()[0]).real
    if normalizer is not None:
        return measure.variance[0] * cov / normalizer
    return cov


def rff_sample(space, measure, lmd, shift, w_re, w_im, x, normalizer):
    """prior_approximation.py:111-117."""
    f = space.features(space.to_group(x), lmd, shift) * torch.sqrt(torch.abs(measure.variance[0]))
    s_re = torch.einsum("nm,m->n", f.real, w_re)
    s_im = torch.einsum("nm,m->n", f.imag, w_im)
    return (s_re - s_im) / math.sqrt(normalizer)


# =====================================================================================================================
# GP regression step (examples/gpr_model.py:12-92, examples/gpr_experiments.py:75-87)
# ---------------------------------------------------------------------------------------------------------------------
# The reference delegates this to gpytorch ~1.7 (pyproject.toml:16-23; absent from the image, PARITY UNPINNED for this
# block): ExactGP + GaussianLikelihood + ExactMarginalLogLikelihood.  Its published algorithm for dense covariances of
# this size is a Cholesky factorisation (psd_safe_cholesky -> LAPACK dpotrf) followed by triangular solves:
#     log p(y) = -1/2 ( r^T S^{-1} r + log det S + n log 2 pi ),  S = K + 1e-6 I + noise I,  r = y - mean,
#     mll = log p(y) / n                                   (ExactMarginalLogLikelihood divides by the number of data)
#     posterior: mean* = m + K*^T S^{-1} r,  cov* = K** + 1e-6 I - K*^T S^{-1} K*
#     noise = softplus(raw_noise) + 1e-4   (GaussianLikelihood's default GreaterThan(1e-4) constraint, raw_noise = 0)
# restated here with torch CPU / LAPACK.
# =====================================================================================================================
def gp_noise(raw_noise):
    return torch.nn.functional.softplus(raw_noise) + 1e-4


def gp_log_marginal(K, noise, mean, y):
    """log p(y) / n for y ~ N(mean, K + 1e-6 I + noise I); differentiable (torch autograd through LAPACK)."""
    n = y.shape[0]
    S = K + (1e-6 + noise) * torch.eye(n, dtype=torch.float64)
    L = torch.linalg.cholesky(S)
    r = (y - mean)[:, None]
    alpha = torch.cholesky_solve(r, L)
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    return -0.5 * ((r * alpha).sum() + logdet + n * math.log(2.0 * math.pi)) / n


def gp_posterior(K, K_star, K_ss, noise, mean, y):
    """Posterior mean and covariance at the test points (K_star: [n, n*], K_ss: [n*, n*])."""
    n = y.shape[0]
    S = K + (1e-6 + noise) * torch.eye(n, dtype=torch.float64)
    L = torch.linalg.cholesky(S)
    alpha = torch.cholesky_solve((y - mean)[:, None], L)[:, 0]
    v = torch.linalg.solve_triangular(L, K_star, upper=False)
    return mean + K_star.T @ alpha, K_ss + 1e-6 * torch.eye(K_ss.shape[0], dtype=torch.float64) - v.T @ v
