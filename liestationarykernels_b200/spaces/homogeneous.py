"""Compact homogeneous spaces M = SO(n)/H — drop-in for `space.py:88-143`, `spaces/stiefel.py`, `spaces/grassmannian.py`.

Points are orthonormal n x m frames.  Kernels and feature maps on M are H-averaged characters of SO(n):
    phi_l(x, y) = d_l * mean_k chi_l( lift(x)^T lift(y) h_k^T ),
with the `average_order` subgroup samples h_k drawn once in the constructor (space.py:97).  The lift and the
pair/average/polynomial evaluation are CUDA kernels (csrc/lsk_homog.cu)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib, ops
from ..space import AbstractManifold, LBEigenspace, cuda_device
from ..utils import hook_content_formula
from .so import SO

dtype = torch.float64


class AveragedLieGroupCharacter(torch.nn.Module):
    """d_l * chi_l averaged over the subgroup samples (space.py:265-281)."""

    def __init__(self, representation, space, chi):
        super().__init__()
        self.representation = representation
        self.space = space
        self.chi = chi

    def forward(self, gammas_x_h):
        vals = self.chi(gammas_x_h).reshape(-1, self.space.average_order)
        return torch.mean(vals, dim=-1)


class AveragedLBEigenspace(LBEigenspace):
    """space.py:221-240"""

    def __init__(self, representation, manifold):
        super().__init__(representation.index, representation.dimension, representation.lb_eigenvalue,
                         representation._position, manifold)
        self.initial_representation = representation
        self.inv_dimension = manifold.compute_inv_dimension(representation.index)

    def compute_phase_function(self):
        return AveragedLieGroupCharacter(self, self.manifold, self.initial_representation.compute_phase_function())


class CompactHomogeneousSpace(AbstractManifold):
    """space.py:88-143"""

    def __init__(self, g, h, average_order):
        AbstractManifold.__init__(self)
        if g.n > 5:
            # lsk_homog_pairs is instantiated for SO(3), SO(4), SO(5) (csrc/lsk_homog.cu); the reference's tables would allow
            # SO(6..8) quotients as well, at O(n^3 A) work per pair
            raise ValueError("homogeneous spaces SO(n)/H are built for n <= 5 (got n = %d)" % g.n)
        self.g, self.h = g, h
        self.dim = self.g.dim - self.h.dim
        self.order, self.average_order = self.g.order, average_order
        self.h_samples = self.sample_H(self.average_order)
        self.lb_eigenspaces = [AveragedLBEigenspace(rep, self) for rep in self.g.lb_eigenspaces]
        # the pair kernel forms tr g and e_2(g) itself: same polynomials, classical (non-spinor) variable map
        from ..classfn import group_plan
        self.plan = group_plan("SO", self.g.n, self.g.order, False)

    @property
    def _lambdas(self):
        return self.g._lambdas

    @property
    def _dims(self):
        return self.g._dims

    def H_to_G(self, h):
        raise NotImplementedError

    def sample_H(self, n):
        return self.H_to_G(self.h.rand(n))

    def rand(self, n=1):
        return self.G_to_M(self.g.rand(n))

    def G_to_M(self, g):
        return g[:, :, :self.m]

    def M_to_G(self, x):
        """Lift of frames to SO(n) with the reference's conventions (stiefel.py:38-48); CUDA kernel `lsk_lift_frames`."""
        _lib.require_cuda()
        lib = _lib.load()
        if x.dim() != 3 or x.shape[1] != self.n or x.shape[2] != self.m:
            raise ValueError("expected [num, %d, %d] frames, got %s" % (self.n, self.m, tuple(x.shape)))
        xc = x.to(dtype).contiguous()                       # G8: rand() returns a strided view
        out = torch.empty((xc.shape[0], self.n, self.n), dtype=dtype, device=xc.device)
        bad = torch.zeros(1, dtype=torch.int32, device=xc.device)
        _lib.check(lib.lsk_lift_frames(_lib.ptr(xc), xc.shape[0], self.n, self.m, _lib.ptr(out), _lib.ptr(bad),
                                       _lib.stream_ptr()), "lsk_lift_frames")
        _lib.count_launch()
        assert int(bad.item()) == 0, "lift does not reproduce the frame (stiefel.py:47)"
        return out

    # -- API-compatibility paths -----------------------------------------------------------------------------
    def pairwise_diff(self, x, y):
        return self.g.pairwise_diff(self.M_to_G(x), self.M_to_G(y), reverse=True)

    def pairwise_embed(self, x, y):
        return self.g.pairwise_embed(self.pairwise_diff(x, y), self.h_samples)

    def dist(self, x, y):
        raise NotImplementedError

    def compute_inv_dimension(self, signature):
        raise NotImplementedError

    # -- fused paths -------------------------------------------------------------------------------------------
    def _pairs(self, gx, gy, ep, n_rows, n_cols, out, ld):
        lib = _lib.load()
        h = self.h_samples.to(dtype).contiguous()
        ep.reserved = (ep.reserved & ~0xf0) | (self._identity_block(h) << 4)
        _lib.check(lib.lsk_homog_pairs(_lib.ptr(gx), _lib.ptr(gy), _lib.ptr(h), n_rows, n_cols, self.n, h.shape[0],
                                       ctypes.byref(ep), ctypes.c_void_p(out), ld, _lib.stream_ptr()), "lsk_homog_pairs")
        _lib.count_launch()

    def _identity_block(self, h):
        """Structure code of the subgroup samples for lsk_homog_pairs (bits 4-7 of `reserved`): the largest t in {3, 2} such that
        every sample is blockdiag(I_t, R) - true for Stiefel manifolds, whose H = SO(n - m) acts on the last n - m columns
        (stiefel.py:25-36) - else 0; 11 when t = 3 and every R is a plane rotation [[c, s], [-s, c]] (V(5,3): H = SO(2), so.py:62-70;
        V(5,4): R = I), which selects the trigonometric-polynomial kernel.  Checked on the tensor itself (h_samples is a plain
        attribute that callers may replace), once per tensor version."""
        key = (h.data_ptr(), h._version, tuple(h.shape))
        if getattr(self, "_hblock_key", None) != key:
            t_found = 0
            if self.n == 5:
                eye = torch.eye(self.n, dtype=dtype, device=h.device)
                for t in (3, 2):
                    ok = torch.equal(h[:, :t, :], eye[:t, :].expand(h.shape[0], t, self.n)) and \
                        torch.equal(h[:, t:, :t], torch.zeros_like(h[:, t:, :t]))
                    if ok:
                        t_found = t
                        break
                if t_found == 3:
                    c, s = h[:, 3, 3], h[:, 3, 4]
                    rot = torch.equal(h[:, 4, 4], c) and torch.equal(h[:, 4, 3], -s) and \
                        bool((torch.abs(c * c + s * s - 1.0) < 1e-14).all())
                    if rot:
                        t_found = 11
            self._hblock_key, self._hblock = key, t_found
        return self._hblock

    def _phase_features(self, x, phases, row_scales, panel):
        """Phi[n, l*P + p] = row_scales[l] * d_l * mean_k Re chi_l(lift(phase_p)^T lift(x_n) h_k^T)."""
        scales = np.asarray(row_scales, dtype=np.float64) * self._dims
        N, P, L = x.shape[0], phases.shape[0], len(self.plan.irreps)
        if panel and P % 4:
            return ops.PanelMatrix.from_dense(self._phase_features(x, phases, row_scales, False))
        gx, gp = self.M_to_G(x), self.M_to_G(phases)
        if panel:
            target = ops.PanelMatrix.for_full_write(N, L * P, x.device)
            buf, ld = target.buf, 0
        else:
            target = torch.zeros((N, L * P), dtype=dtype, device=x.device)
            buf, ld = target, target.stride(0)
        for ep, col_shift in self.plan.feature_epilogues(scales, P, monomial_list=True):
            ep.reserved = 1                                   # pair (phase, point): M = lift(phase)^T lift(x)
            if panel:
                ep.out_layout, ep.out_kg = _lib.OUT_PANEL_MAJOR, target.kg
                base = buf.data_ptr() + (col_shift // 4) * 256 * 8
            else:
                ep.out_layout = _lib.OUT_ROW_MAJOR
                base = buf.data_ptr() + col_shift * 8
            if N and P:
                self._pairs(gx, gp, ep, N, P, base, ld)
        return target

    def _phase_sample(self, x, phases, row_scales, weights):
        """f[n] = sum_l sum_p weights[l * P + p] * Phi[n, l * P + p] with Phi = `_phase_features(x, phases, row_scales)`: the
        prior sample of RandomPhaseApproximation (prior_approximation.py:85-88).  On SO(5)/SO(2) the feature kernel reduces over
        the phases itself (`lsk_homog_sample`: the [N, L * P] map is never written); otherwise the map is written panel-major and
        read once by `lsk_panel_gemv`."""
        h = self.h_samples.to(dtype).contiguous()
        N, P = x.shape[0], phases.shape[0]
        if self.n == 5 and self._identity_block(h) == 11 and N and P:
            lib = _lib.load()
            scales = np.asarray(row_scales, dtype=np.float64) * self._dims
            batches = list(self.plan.feature_epilogues(scales, P, monomial_list=True))
            if len(batches) == 1 and batches[0][1] == 0:
                ep = batches[0][0]
                ep.reserved = 1 | (11 << 4)                  # pair (phase, point); rotation samples
                gx, gp = self.M_to_G(x), self.M_to_G(phases)
                w = weights.to(dtype).contiguous()
                out = torch.empty(N, dtype=dtype, device=x.device)
                rc = lib.lsk_homog_sample(_lib.ptr(gx), _lib.ptr(gp), _lib.ptr(h), N, P, self.n, h.shape[0], ctypes.byref(ep),
                                          _lib.ptr(w), _lib.ptr(out), _lib.stream_ptr())
                if rc == 0:
                    _lib.count_launch()
                    return out
                if rc != -3:                                  # LSK_EUNSUPPORTED falls through to the two-pass path
                    _lib.check(rc, "lsk_homog_sample")
        return ops.panel_matvec(self._phase_features(x, phases, row_scales, panel=True), weights)

    def _eigensum(self, kernel, x, y, normalize, out=None):
        """EigenbasisSumKernel on M (spectral_kernel.py:37-54 with space.py:131-135, 273-275)."""
        from ..spectral_kernel import spectral_weights
        w = spectral_weights(kernel.measure, self)
        scale = float(torch.abs(kernel.measure.variance[0]).item() / float(kernel.normalizer)) if normalize else 1.0
        return self._eigensum_coefs(x, y, w, scale, out)

    def _identity_values(self):
        """phi_l(id, id) = d_l mean_k Re chi_l(h_k^T) for every irrep: the value each irrep contributes to the kernel at
        coinciding points, i.e. to the normaliser (spectral_kernel.py:33-35).  Fixed by the subgroup samples; cached."""
        key = (self.h_samples.data_ptr(), self.h_samples._version)
        if getattr(self, "_identity_cache", (None, None))[0] != key:
            eye = torch.zeros((1, self.n, self.m), dtype=dtype, device=self.h_samples.device)
            eye[0, :self.m, :self.m] = torch.eye(self.m, dtype=dtype, device=eye.device)
            vals = self._phase_features(eye, eye, np.ones(len(self.plan.irreps)), False).reshape(-1)
            self._identity_cache = (key, vals)
        return self._identity_cache[1]

    def _eigensum_coefs(self, x, y, coefs, scale, out=None):
        """scale * sum_l coefs[l] phi_l(x_i, y_j) for explicit per-irrep coefficients (the differentiable kernel path
        evaluates this once with c(theta) and once per hyper-parameter with dc/dtheta)."""
        w = np.asarray(coefs, dtype=np.float64) * self._dims
        ep = self.plan.collapsed_row_epilogue(w, scale)
        ep.out_layout = _lib.OUT_ROW_MAJOR
        gx = self.M_to_G(x)
        gy = gx if y is x else self.M_to_G(y)
        if out is None:
            out = torch.empty((x.shape[0], y.shape[0]), dtype=dtype, device=x.device)
        if x.shape[0] and y.shape[0]:
            self._pairs(gx, gy, ep, x.shape[0], y.shape[0], out.data_ptr(), out.stride(0))
        return out


def _block_diag(hu, hd):
    b, nu, nd = hu.shape[0], hu.shape[1], hd.shape[1]
    out = torch.zeros((b, nu + nd, nu + nd), dtype=dtype, device=hu.device)
    out[:, :nu, :nu] = hu
    out[:, nu:, nu:] = hd
    return out


class Stiefel(CompactHomogeneousSpace):
    """V(n, m) = SO(n)/SO(n-m), points are orthonormal n x m frames (stiefel.py:11-69)."""

    def __init__(self, n, m, order=10, average_order=30):
        assert n > m, "n should be greater than m"
        self.n, self.m = n, m
        self.n_m = n - m
        g = SO(self.n, order=order)
        h = SO(self.n_m, order=0)
        CompactHomogeneousSpace.__init__(self, g=g, h=h, average_order=average_order)
        self.id = torch.zeros((self.n, self.m), device=cuda_device(), dtype=dtype).fill_diagonal_(1.0).view((1, self.n, self.m))

    def H_to_G(self, h):
        eye = torch.eye(self.m, device=h.device, dtype=dtype).view(1, self.m, self.m).repeat(h.size()[0], 1, 1)
        return _block_diag(eye, h)

    def compute_inv_dimension(self, signature):
        m_ = min(self.m, self.n_m)
        if m_ < self.g.rank and signature[m_] > 0:
            return 0
        return hook_content_formula(tuple(abs(x) for x in signature), m_)


class _SOxSO:
    """SO(n) x SO(m), block diagonal (grassmannian.py:12-27)."""

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.so_n, self.so_m = SO(n, order=0), SO(m, order=0)
        self.dim = n * (n - 1) // 2 + m * (m - 1) // 2

    def rand(self, n=1):
        return _block_diag(self.so_n.rand(n), self.so_m.rand(n))


class _S_OxO:
    """S(O(n) x O(m)) (grassmannian.py:29-55)."""

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.so_n, self.so_m = SO(n, order=0), SO(m, order=0)
        self.dim = n * (n - 1) // 2 + m * (m - 1) // 2 - 1

    def rand(self, n=1):
        assert n % 2 == 0, "number of samples must be even."
        hu, hd = self.so_n.rand(n // 2), self.so_m.rand(n // 2)
        first = _block_diag(hu, hd).clone()
        hu[:, :, -1] = hu[:, :, -1] * -1
        hd[:, :, -1] = hd[:, :, -1] * -1
        return torch.cat((first, _block_diag(hu, hd)), dim=0)


class OrientedGrassmannian(CompactHomogeneousSpace):
    """SO(n)/(SO(m) x SO(n-m))  (grassmannian.py:57-102)."""

    def __init__(self, n, m, order=10, average_order=50):
        assert n > m, "n should be greater than m"
        self.n, self.m = n, m
        self.n_m = n - m
        g = SO(n, order=order)
        h = _SOxSO(self.m, self.n_m)
        CompactHomogeneousSpace.__init__(self, g=g, h=h, average_order=average_order)
        self.id = torch.zeros((self.n, self.m), device=cuda_device(), dtype=dtype).fill_diagonal_(1.0).view((1, self.n, self.m))

    def H_to_G(self, h):
        return h

    def compute_inv_dimension(self, signature):
        return 1


class Grassmannian(CompactHomogeneousSpace):
    """SO(n)/S(O(m) x O(n-m))  (grassmannian.py:103-155)."""

    def __init__(self, n, m, order=10, average_order=30):
        assert n > m, "n should be greater than m"
        self.n, self.m = n, m
        self.n_m = n - m
        g = SO(n, order=order)
        h = _S_OxO(self.m, self.n_m)
        CompactHomogeneousSpace.__init__(self, g=g, h=h, average_order=average_order)
        self.id = torch.zeros((self.n, self.m), device=cuda_device(), dtype=dtype).fill_diagonal_(1.0).view((1, self.n, self.m))

    def H_to_G(self, h):
        return h

    def dist(self, x, y):
        _, s, _ = torch.linalg.svd(torch.bmm(torch.transpose(x, -2, -1), y))
        s = torch.clamp(s, max=1.0)
        return torch.linalg.norm(torch.arccos(s), dim=1)

    def pairwise_dist(self, x, y):
        nx, ny = x.shape[0], y.shape[0]
        xf = x[:, None].expand(nx, ny, self.n, self.m).reshape(-1, self.n, self.m)
        yf = y[None, :].expand(nx, ny, self.n, self.m).reshape(-1, self.n, self.m)
        return self.dist(xf, yf).reshape(nx, ny)

    def compute_inv_dimension(self, signature):
        return 1
