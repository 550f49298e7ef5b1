"""Non-compact symmetric spaces: hyperbolic space H^n (ball model) and SPD(n) = GL(n)/O(n) — drop-in for
`spaces/hyperbolic.py` and `spaces/spd.py` on the random-Fourier-feature path.

Spectral sampling (`generate_lb_eigenspaces`) keeps the reference's torch RNG calls in the reference's order, so
seeded scripts draw the same frequencies and shifts; the parity quirks G6/G9 of SURVEY.md are preserved (H^n Matern
scale uses 1/4 for every n and caches the normalised frequencies; SPD Matern scale uses l, not l^2, and re-samples
on every call; pi is the float32-rounded 3.1415927410125732 the reference builds from a float32 acos).
The N x m feature evaluation is a fused CUDA kernel (csrc/lsk_noncompact.cu)."""
from __future__ import annotations

from math import factorial, sqrt

import torch
from torch.distributions import Normal, StudentT

from .. import _lib, ops
from ..space import NonCompactSymmetricSpace, cuda_device
from ..spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure

dtype = torch.float64
REF_PI = 3.1415927410125732          # 2 * float32(acos(0)), hyperbolic.py:11 / spd.py:12


class _SphericalFunctions(torch.nn.Module):
    """Callable feature map with the reference's attribute layout: `.exp.lmd`, `.exp.shift`, `.coeff`
    (HypShiftedNormailizedExp / SPDShiftedNormailizedExp)."""

    class _Exp:
        def __init__(self, lmd, shift, manifold):
            self.lmd, self.shift, self.manifold = lmd, shift, manifold
            self.order = lmd.size()[0]

    def __init__(self, lmd, shift, manifold, coeff):
        super().__init__()
        self.n = manifold.n
        self.manifold = manifold
        self.exp = self._Exp(lmd, shift, manifold)
        self.coeff = coeff

    def forward(self, x_group):
        """complex128 [N, m] features of points given in group form (`manifold.to_group(x)`)."""
        return self.manifold._features(x_group, self, scale=1.0, layout="complex")


class HyperbolicSpace(NonCompactSymmetricSpace):
    """Ball model of H^n (hyperbolic.py:14-105)."""

    def __init__(self, n: int, order=10000):
        super().__init__()
        self.n = n
        self.dim = n
        self.order = order
        self.id = torch.zeros(self.n, device=cuda_device(), dtype=dtype).view(1, self.n)
        self.normalized_lmd = None

    def _generate_lb_eigenspace(self, measure):
        dev = cuda_device()
        if isinstance(measure, MaternSpectralMeasure):
            df = 2 * measure.nu + measure.dim
            student = StudentT(df=df - 1, scale=1).rsample((self.order,))
            lmd = torch.abs(student) / torch.sqrt(df - 1)
        elif isinstance(measure, SqExpSpectralMeasure):
            normal = Normal(torch.tensor(0, dtype=dtype, device=dev), torch.tensor(1, dtype=dtype, device=dev)).rsample((self.order,)).type(dtype)
            lmd = torch.abs(normal)
        else:
            return NotImplementedError
        self.normalized_lmd = torch.squeeze(lmd)
        self.shift = self.rand_phase(self.order)

    def generate_lb_eigenspaces(self, measure):
        if self.normalized_lmd is None:
            self._generate_lb_eigenspace(measure)
        if isinstance(measure, MaternSpectralMeasure):
            scale = torch.sqrt(1 / 4 + 2 * measure.nu[0] / (measure.lengthscale[0] ** 2))
        elif isinstance(measure, SqExpSpectralMeasure):
            scale = 1.0 / torch.abs(measure.lengthscale[0])
        else:
            return NotImplementedError
        lmd = (self.normalized_lmd * scale).detach()
        self.lb_eigenspaces = _SphericalFunctions(lmd, self.shift, self, self.c_function(lmd))

    def _differentiable_spectrum(self, measure, resample):
        """(lmd, shift, coeff) with lmd and coeff attached to the autograd graph of the lengthscale (the normalised samples
        are drawn once and cached, hyperbolic.py:27-52; eval mode keeps the features of the last generation)."""
        if self.normalized_lmd is None:
            self._generate_lb_eigenspace(measure)
        if isinstance(measure, MaternSpectralMeasure):
            scale = torch.sqrt(1 / 4 + 2 * measure.nu[0] / (measure.lengthscale[0] ** 2))
        else:
            scale = 1.0 / torch.abs(measure.lengthscale[0])
        if resample:
            lmd, shift = self.normalized_lmd.detach() * scale, self.shift
        else:
            lmd, shift = self.lb_eigenspaces.exp.lmd.detach() * (scale / scale.detach()), self.lb_eigenspaces.exp.shift
        return lmd, shift, self.c_function(lmd)

    def _log_terms(self, x, funcs, grouped=False):
        """L[i, k] = log((1 - |x_i|^2) / |x_i - b_k|^2): dF_ik / dlmd_k = -i L_ik F_ik"""
        pg, shift = (x if grouped else self.to_group(x)), funcs.exp.shift
        d2 = torch.sum(torch.square(pg[:, None, :] - shift[None, :, :]), dim=-1)
        return -(torch.log(1 - torch.sum(torch.square(pg), dim=1))[:, None] - torch.log(d2))[:, :, None]

    def c_function(self, lmd):
        """|c(lambda)|^{-1} up to constants (hyperbolic.py:134-150)."""
        n = self.n
        if n % 2 == 0:
            adds = torch.tensor([(2 * i + 1) ** 2 / 4 for i in range(n // 2 - 1)], dtype=dtype, device=lmd.device)
        else:
            adds = torch.tensor([i ** 2 for i in range(n // 2)], dtype=dtype, device=lmd.device)
        log_c = torch.sum(torch.log(torch.square(lmd)[:, None] + adds[None, :]), dim=1)
        if n % 2 == 0:
            log_c = log_c + torch.log(lmd) + torch.log(torch.tanh(REF_PI * lmd))
        return torch.squeeze(torch.exp(log_c / 2))

    def to_group(self, x):
        return x.squeeze(dim=-1)

    def rand_phase(self, n=1):
        if n == 0:
            return None
        x = torch.randn(n, self.n, device=cuda_device(), dtype=dtype)
        return x / torch.norm(x, dim=1, keepdim=True)

    def rand(self, n=1):
        sphere = self.rand_phase(n)
        r = torch.pow(torch.rand(n, device=cuda_device(), dtype=dtype), 1 / self.n) * 0.99
        return sphere * r[:, None].clone()

    def inv(self, x):
        return -x

    def _dist_to_id(self, x):
        e = torch.sqrt(torch.sum(torch.square(x), dim=1))
        return torch.log((1 + e) / (1 - e))

    def pairwise_dist(self, x, y):
        d2 = torch.sum(torch.square(x[:, None] - y[None, :]), dim=-1)
        nx, ny = torch.sum(torch.square(x), dim=1), torch.sum(torch.square(y), dim=1)
        return torch.arccosh(1 + 2 * d2 / (1 - nx)[:, None] / (1 - ny)[None, :])

    def pairwise_diff(self, x, y):
        """hyperbolic.py:57-69: the point at hyperbolic distance d(x_i, y_j) from the origin on the diagonal ray."""
        dist = self.pairwise_dist(x, y).reshape(-1)
        ones = torch.ones((dist.size()[0], self.n), device=x.device, dtype=dtype) / sqrt(self.n)
        return ones * torch.tanh(dist / 2)[:, None]

    def _features(self, x_group, funcs, scale, layout, out=None):
        return _run_features(self, "hyp", x_group.to(dtype).contiguous(), funcs, scale, layout)

    def _sample(self, x_group, funcs, scale, weights):
        return _run_sample(self, "hyp", x_group.to(dtype).contiguous(), funcs, scale, weights)


class SymmetricPositiveDefiniteMatrices(NonCompactSymmetricSpace):
    """SPD(n) = GL(n, R)/O(n, R)  (spd.py:15-80)."""

    def __init__(self, n: int, order=200):
        super().__init__()
        self.n = n
        self.dim = n * (n + 1) // 2
        self.order = order
        self.id = torch.eye(self.n, device=cuda_device(), dtype=dtype).view(1, self.n, self.n)

    def _draw(self, measure):
        """The random part of spd.py:27-43 in the reference's RNG order: shifts, GOE spectra, chi^2 mixing (Matern only)."""
        from ..utils import GOE_sampler
        shift = self.rand_phase(self.order)
        goe = GOE_sampler(self.order, self.n)
        chi2 = None
        if isinstance(measure, MaternSpectralMeasure):
            chi2 = torch.distributions.chi2.Chi2(2 * measure.nu[0]).rsample((self.order,))
        return shift, goe, chi2

    def _inverse_scale(self, measure):
        """lmd = base * this: sqrt((n^3 - n) / 48 + 2 nu / l) for Matern (spd.py:31: l, not l^2), 1 / l for the heat kernel."""
        if isinstance(measure, MaternSpectralMeasure):
            return torch.sqrt((self.n ** 3 - self.n) / 48 + 2 * measure.nu[0] / measure.lengthscale[0])
        return 1.0 / measure.lengthscale[0]

    def generate_lb_eigenspaces(self, measure):
        if not isinstance(measure, (MaternSpectralMeasure, SqExpSpectralMeasure)):
            return NotImplementedError
        shift, goe, chi2 = self._draw(measure)
        if chi2 is not None:
            scale = 1 / torch.sqrt((self.n ** 3 - self.n) / 48 + 2 * measure.nu[0] / measure.lengthscale[0])
            lmd = goe / torch.sqrt(chi2)[:, None] / scale
        else:
            lmd = goe / measure.lengthscale
        lmd = lmd.detach()
        self._base_lmd = None
        self.lb_eigenspaces = _SphericalFunctions(lmd, shift, self, self.c_function_tanh(lmd))

    def _differentiable_spectrum(self, measure, resample):
        """(lmd [m, n], shift, coeff) with lmd and coeff attached to the autograd graph of the lengthscale.  resample: new
        shifts and spectra as `generate_lb_eigenspaces` draws them on every training-mode forward (spd.py:28)."""
        inv_scale = self._inverse_scale(measure)
        if resample:
            shift, goe, chi2 = self._draw(measure)
            base = goe if chi2 is None else goe / torch.sqrt(chi2)[:, None]
            lmd = base.detach() * inv_scale
        else:
            shift = self.lb_eigenspaces.exp.shift
            lmd = self.lb_eigenspaces.exp.lmd.detach() * (inv_scale / inv_scale.detach())
        return lmd, shift, self.c_function_tanh(lmd)

    def _log_terms(self, x, funcs, grouped=False):
        """la[i, k, r] = log a_r(U_i h_k^{-1}): dF_ik / dlmd_kr = i la_r F_ik  (CUDA kernel `lsk_spd_logdiag`)"""
        lib = _lib.load()
        u = (x if grouped else self.to_group(x)).to(dtype).contiguous()
        hinv = self.inv(funcs.exp.shift)
        m = funcs.exp.lmd.shape[0]
        out = torch.empty((u.shape[0], m, self.n), dtype=dtype, device=u.device)
        if u.shape[0] and m:
            _lib.check(lib.lsk_spd_logdiag(_lib.ptr(u), _lib.ptr(hinv), u.shape[0], m, self.n, _lib.ptr(out), _lib.stream_ptr()),
                       "lsk_spd_logdiag")
            _lib.count_launch()
        return out

    def c_function_tanh(self, lmd):
        """spd.py:116-123"""
        iu = torch.triu_indices(self.n, self.n, offset=1, device=lmd.device)
        diff = (lmd[:, None, :] - lmd[:, :, None])[:, iu[0], iu[1]]
        return torch.exp(torch.sum(torch.log(torch.tanh(REF_PI * torch.abs(diff))), dim=1) / 2)

    def to_group(self, x):
        """Upper Cholesky factor (spd.py:45-46); CUDA kernel `lsk_small_matrix`."""
        return _small_matrix(x, 0, "matrix is not positive definite")

    def inv(self, x):
        return _small_matrix(x, 1, "matrix is singular")

    def rand_phase(self, n=1):
        from .so import haar_from_gaussian
        g = torch.randn((n, self.n, self.n), device=cuda_device(), dtype=dtype)
        return haar_from_gaussian(g)

    def rand(self, n=1):
        g = torch.randn(n, self.n, self.n, device=cuda_device(), dtype=dtype)
        return torch.bmm(g, torch.transpose(g, -2, -1)) * 4 / (factorial(self.n + 1) ** (1 / self.n))

    def _dist_to_id(self, x):
        eig = torch.linalg.eigvalsh(x)
        return torch.sum(torch.square(torch.log(eig)), dim=1)

    def pairwise_diff(self, x, y):
        """space.py:338-346: x_i y_j^{-1}, flattened [N*M, n, n]."""
        return torch.matmul(x[:, None], self.inv(y)[None, :]).reshape(-1, self.n, self.n)

    def pairwise_dist(self, x, y):
        xy = torch.matmul(self.to_group(x)[:, None], self.inv(self.to_group(y))[None, :]).reshape(-1, self.n, self.n)
        return self._dist_to_id(xy).reshape(x.shape[0], y.shape[0])

    def _features(self, x_group, funcs, scale, layout, out=None):
        return _run_features(self, "spd", x_group.to(dtype).contiguous(), funcs, scale, layout)

    def _sample(self, x_group, funcs, scale, weights):
        return _run_sample(self, "spd", x_group.to(dtype).contiguous(), funcs, scale, weights)


def _small_matrix(x, mode, msg):
    _lib.require_cuda()
    lib = _lib.load()
    n = x.shape[-1]
    xc = x.to(dtype).contiguous().reshape(-1, n, n)
    out = torch.empty_like(xc)
    bad = torch.zeros(1, dtype=torch.int32, device=xc.device)
    _lib.check(lib.lsk_small_matrix(_lib.ptr(xc), xc.shape[0], n, mode, _lib.ptr(out), _lib.ptr(bad), _lib.stream_ptr()),
               "lsk_small_matrix")
    _lib.count_launch()
    if int(bad.item()):
        raise RuntimeError("linalg: %s" % msg)
    return out.reshape(x.shape)


def _run_sample(space, kind, x_group, funcs, scale, weights):
    """f[i] = scale * sum_k (Re F[i,k] weights[k] + Im F[i,k] weights[m + k]) with the feature map fused away
    (`lsk_hyp_sample` / `lsk_spd_sample`)."""
    _lib.require_cuda()
    lib = _lib.load()
    lmd, shift, coeff = funcs.exp.lmd, funcs.exp.shift, funcs.coeff
    m, N = lmd.shape[0], x_group.shape[0]
    if weights.numel() != 2 * m:
        raise ValueError("expected %d weights, got %d" % (2 * m, weights.numel()))
    lmd_c = lmd.to(dtype).contiguous()
    coef_c = coeff.to(dtype).reshape(-1).contiguous()
    w = weights.to(dtype).contiguous()
    out = torch.empty(N, dtype=dtype, device=x_group.device)
    if N == 0:
        return out
    if kind == "hyp":
        sh = shift.to(dtype).contiguous()
        _lib.check(lib.lsk_hyp_sample(_lib.ptr(x_group), _lib.ptr(sh), _lib.ptr(lmd_c), _lib.ptr(coef_c), _lib.ptr(w), N, m,
                                      space.n, float(scale), _lib.ptr(out), _lib.stream_ptr()), "lsk_hyp_sample")
    else:
        hinv = space.inv(shift)
        _lib.check(lib.lsk_spd_sample(_lib.ptr(x_group), _lib.ptr(hinv), _lib.ptr(lmd_c), _lib.ptr(coef_c), _lib.ptr(w), N, m,
                                      space.n, float(scale), _lib.ptr(out), _lib.stream_ptr()), "lsk_spd_sample")
    _lib.count_launch()
    return out


def _run_features(space, kind, x_group, funcs, scale, layout):
    """layout: 'complex' -> complex128 [N, m];  'dense' -> float64 [N, 2m] = [Re | Im];  'panel' -> ops.PanelMatrix."""
    _lib.require_cuda()
    lib = _lib.load()
    lmd, shift, coeff = funcs.exp.lmd, funcs.exp.shift, funcs.coeff
    m, N = lmd.shape[0], x_group.shape[0]
    lmd_c = lmd.to(dtype).contiguous()
    coef_c = coeff.to(dtype).reshape(-1).contiguous()
    if layout == "panel" and m % 4:
        return ops.PanelMatrix.from_dense(_run_features(space, kind, x_group, funcs, scale, "dense"))
    if layout == "complex":
        target = torch.empty((N, m), dtype=torch.complex128, device=x_group.device)
        buf, code, ld, kg = torch.view_as_real(target), _lib.FEAT_COMPLEX, m, 0
    elif layout == "dense":
        target = torch.empty((N, 2 * m), dtype=dtype, device=x_group.device)
        buf, code, ld, kg = target, _lib.OUT_ROW_MAJOR, 2 * m, 0
    else:
        target = ops.PanelMatrix.for_full_write(N, 2 * m, x_group.device)
        buf, code, ld, kg = target.buf, _lib.OUT_PANEL_MAJOR, 0, target.kg
    if kind == "hyp":
        sh = shift.to(dtype).contiguous()
        _lib.check(lib.lsk_hyp_features(_lib.ptr(x_group), _lib.ptr(sh), _lib.ptr(lmd_c), _lib.ptr(coef_c), N, m, space.n,
                                        float(scale), _lib.ptr(buf), code, ld, kg, _lib.stream_ptr()), "lsk_hyp_features")
    else:
        hinv = space.inv(shift)
        _lib.check(lib.lsk_spd_features(_lib.ptr(x_group), _lib.ptr(hinv), _lib.ptr(lmd_c), _lib.ptr(coef_c), N, m, space.n,
                                        float(scale), _lib.ptr(buf), code, ld, kg, _lib.stream_ptr()), "lsk_spd_features")
    _lib.count_launch()
    return target
