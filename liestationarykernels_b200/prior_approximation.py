"""Prior samplers — drop-in for the reference's `prior_approximation.py:50-123`
(`RandomPhaseApproximation`, `RandomFourierApproximation`; same attributes `weights`, `phases`, `weights_real`,
`weights_imag`, same RNG call order in constructors and `resample()`)."""
from __future__ import annotations

from math import sqrt

import numpy as np
import torch

from . import ops
from .space import cuda_device
from .spectral_kernel import spectral_weights

dtype = torch.float64


class RandomPhaseApproximation(torch.nn.Module):
    """f(x) = Phi(x) w,  w ~ N(0, I_{L*P}),  Phi scaled by sqrt(|sigma^2| a_l / norm / P)  (prior_approximation.py:50-92)."""

    def __init__(self, kernel, phase_order=1000):
        super().__init__()
        self.kernel = kernel
        self.approx_order = len(self.kernel.manifold.lb_eigenspaces)
        self.phase_order = phase_order
        self.weights = self.sample_weights()
        self.phases = self.sample_phases()

    def sample_weights(self):
        return torch.randn(self.phase_order * self.approx_order, device=cuda_device(), dtype=dtype)

    def sample_phases(self):
        return self.kernel.manifold.rand(self.phase_order)

    def resample(self):
        self.weights = self.sample_weights()
        self.phases = self.sample_phases()

    def _row_scales(self):
        k = self.kernel
        a = spectral_weights(k.measure, k.manifold)
        var = abs(float(k.measure.variance[0]))
        return np.sqrt(var * a) / sqrt(float(k.normalizer)) / sqrt(self.phase_order)

    def make_embedding(self, x):
        return self.kernel.manifold._phase_features(x, self.phases, self._row_scales(), panel=False)

    def forward(self, x):
        m = self.kernel.manifold
        if hasattr(m, "_phase_sample"):        # fused: the [N, L * P] feature map is not written where the kernel can reduce it
            return m._phase_sample(x, self.phases, self._row_scales(), self.weights)
        feats = m._phase_features(x, self.phases, self._row_scales(), panel=True)
        return ops.panel_matvec(feats, self.weights)

    def _cov(self, x, y):
        scales = self._row_scales()
        fx = self.kernel.manifold._phase_features(x, self.phases, scales, panel=True)
        fy = self.kernel.manifold._phase_features(y, self.phases, scales, panel=True)
        return ops.gram(fx, fy)


class RandomFourierApproximation(torch.nn.Module):
    """f(x) = (Re F w_re - Im F w_im) sqrt(|sigma^2|) / sqrt(norm)   (prior_approximation.py:95-123)."""

    def __init__(self, kernel):
        super().__init__()
        self.kernel = kernel
        self.weights_real = self.sample_weights()
        self.weights_imag = self.sample_weights()

    def sample_weights(self):
        return torch.randn(self.kernel.manifold.order, dtype=dtype, device=cuda_device())

    def resample(self):
        self.weights_real = self.sample_weights()
        self.weights_imag = self.sample_weights()
        self.kernel.manifold.generate_lb_eigenspaces(self.kernel.measure)

    def _scale(self):
        k = self.kernel
        return sqrt(abs(float(k.measure.variance[0]))) / sqrt(float(k.normalizer))

    def forward(self, x):
        m = self.kernel.manifold
        w = torch.cat((self.weights_real, -self.weights_imag))
        if hasattr(m, "_sample"):          # fused: the [N, 2m] feature map is never written
            return m._sample(m.to_group(x), m.lb_eigenspaces, self._scale(), w)
        return ops.panel_matvec(m._features(m.to_group(x), m.lb_eigenspaces, self._scale(), "panel"), w)

    def _cov(self, x, y):
        m = self.kernel.manifold
        px = m._features(m.to_group(x), m.lb_eigenspaces, self._scale(), "panel")
        py = m._features(m.to_group(y), m.lb_eigenspaces, self._scale(), "panel")
        return ops.gram(px, py)
