"""Exact GP regression on top of the covariance kernels — the caller of the hot path in the reference
(`examples/gpr_model.py:12-92`, `examples/gpr_experiments.py:75-87`), without gpytorch (absent from the image; SURVEY.md
8f row f2).  Same model and training loop: constant mean, Gaussian likelihood with gpytorch's default noise
parametrisation (`softplus(raw_noise) + 1e-4`), `covar = kernel(x) + 1e-6 I`, exact marginal log-likelihood divided by
the number of targets, Adam + StepLR.

The marginal likelihood and the posterior run on this package's fp64 Cholesky (`linalg.CholeskyFactor` ->
`lsk_potrf_upper` / `lsk_potrs_upper` / `lsk_potri_upper`, csrc/lsk_cholesky.cu).  Hyper-parameter gradients flow through
`EigenbasisSumKernel`'s fused backward (spectral_kernel._WeightedCharacterSum); the gradient of the log-determinant needs
the dense inverse (`CholeskyFactor.inverse`).

Like gpytorch, every call is pure: the diagonal shifts (`+ 1e-6 I`, `+ noise I`) are carried as a pending scalar on the
`MultivariateNormal` and applied inside `log_prob` to the copy of the covariance the factorisation overwrites anyway (no
N x N identity, no extra pass, nothing the caller holds is modified; `likelihood(output)` and `mll(output, y)` can be
evaluated any number of times on the same `output`)."""
from __future__ import annotations

import math

import torch
from torch.optim.lr_scheduler import StepLR

from .linalg import CholeskyFactor

dtype = torch.float64


def _dense(k):
    """dense covariance of a kernel result: a tensor, a LazyGram, or a real gpytorch LazyTensor / linear_operator"""
    if hasattr(k, "evaluate"):
        return k.evaluate()
    return k.to_dense() if hasattr(k, "to_dense") and not isinstance(k, torch.Tensor) else k


class _GaussianLogProb(torch.autograd.Function):
    """log N(resid; 0, covar + shift I) for a dense covar, with gradients
       d/dcovar = (alpha alpha^T - S^{-1}) / 2,  d/dshift = tr(d/dcovar),  d/dresid = -alpha,  S = covar + shift I, alpha = S^{-1} resid.
    The shift is added to the private copy that the factorisation overwrites; `covar` itself is never written."""

    @staticmethod
    def forward(ctx, covar, shift, resid):
        work = covar.detach().clone(memory_format=torch.contiguous_format)
        work.diagonal().add_(shift.detach().reshape(()))
        f = CholeskyFactor(work, overwrite=True)
        alpha = f.solve(resid)
        n = resid.shape[0]
        ctx.factor, ctx.alpha = f, alpha
        return -0.5 * (torch.dot(resid, alpha) + f.logdet() + n * math.log(2.0 * math.pi))

    @staticmethod
    def backward(ctx, g):
        f, alpha = ctx.factor, ctx.alpha
        grad_covar = grad_shift = grad_resid = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gs = float(g)
            grad_covar = f.inverse().addr_(alpha, alpha, beta=-0.5 * gs, alpha=0.5 * gs)     # one pass over the N x N inverse
            if ctx.needs_input_grad[1]:
                grad_shift = torch.diagonal(grad_covar).sum().reshape(1)
            if not ctx.needs_input_grad[0]:
                grad_covar = None
        if ctx.needs_input_grad[2]:
            grad_resid = -g * alpha
        return grad_covar, grad_shift, grad_resid


class MultivariateNormal:
    """The slice of `gpytorch.distributions.MultivariateNormal` the reference's scripts use.  The covariance is
    `base + diag_shift * I`; the shift stays symbolic (a [1] tensor on the autograd graph of the noise) until someone needs
    the dense matrix, so adding jitter or noise never touches the N x N buffer."""

    def __init__(self, mean, covariance_matrix, diag_shift=None):
        self.mean = mean
        self._base = covariance_matrix
        self._shift = diag_shift

    loc = property(lambda self: self.mean)

    def _shift_tensor(self):
        if self._shift is None:
            return torch.zeros(1, dtype=dtype, device=self._base.device)
        return self._shift

    def add_jitter(self, shift):
        """N(mean, covariance + shift I) sharing this distribution's base matrix (read-only)."""
        if not isinstance(shift, torch.Tensor):
            shift = torch.tensor([shift], dtype=dtype, device=self._base.device)
        total = shift.reshape(1) if self._shift is None else self._shift + shift.reshape(1)
        return MultivariateNormal(self.mean, self._base, total)

    @property
    def covariance_matrix(self):
        """Dense covariance; a fresh tensor whenever a shift is pending (the base buffer is never modified)."""
        if self._shift is None:
            return self._base
        out = self._base.clone()
        out.diagonal().add_(self._shift.reshape(()))
        return out

    @property
    def variance(self):
        d = torch.diagonal(self._base)
        return d if self._shift is None else d + self._shift.reshape(())

    @property
    def stddev(self):
        return self.variance.clamp_min(0).sqrt()

    def log_prob(self, value):
        return _GaussianLogProb.apply(self._base, self._shift_tensor(), value - self.mean)


class GaussianLikelihood(torch.nn.Module):
    """Homoskedastic noise, gpytorch's default parametrisation: noise = softplus(raw_noise) + 1e-4, raw_noise = 0."""

    def __init__(self, device=None):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1, dtype=dtype, device=device))

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.raw_noise) + 1e-4

    def forward(self, dist: MultivariateNormal) -> MultivariateNormal:
        return dist.add_jitter(self.noise)


class ConstantMean(torch.nn.Module):
    def __init__(self, device=None):
        super().__init__()
        self.constant = torch.nn.Parameter(torch.zeros(1, dtype=dtype, device=device))

    def forward(self, x):
        return self.constant.expand(x.shape[0])


class ExactGPModel(torch.nn.Module):
    """`examples/gpr_model.py:12-39`.  Training mode: `model(train_x)` is the prior at the training inputs.  Eval mode:
    `model(x)` is the posterior at x given (train_x, train_y) under the likelihood's noise (gpytorch ExactGP semantics)."""

    def __init__(self, train_x, train_y, likelihood, kernel, manifold, point_shape=None):
        super().__init__()
        self.train_inputs, self.train_targets = train_x, train_y
        self.likelihood = likelihood
        self.mean_module = ConstantMean(device=train_x.device)
        self.covar_module = kernel
        self.n = manifold.n
        self.point_shape = point_shape
        self._cache = None

    def _points(self, x):
        if self.point_shape is not None:
            return x.view(*x.shape[:-1], *self.point_shape)
        if torch.is_complex(x):
            return torch.view_as_real(x).reshape(x.shape[0], -1)
        return x

    def forward(self, x):
        mean_x = self.mean_module(x)
        return MultivariateNormal(mean_x, _dense(self.covar_module(self._points(x)))).add_jitter(1e-6)

    def train(self, mode=True):
        self._cache = None
        return super().train(mode)

    def __call__(self, x):
        if self.training:
            return self.forward(x)
        return self._posterior(x)

    def _posterior(self, x):
        tx, ty = self.train_inputs, self.train_targets
        n = tx.shape[0]
        if self._cache is None:
            prior = self.likelihood(self.forward(tx))
            f = CholeskyFactor(prior.covariance_matrix.detach(), overwrite=True)
            self._cache = (f, f.solve((ty - prior.mean).detach()))
        f, alpha = self._cache
        pts, tpts = self._points(x), self._points(tx)
        k_star = _dense(self.covar_module(tpts, pts)).detach()                  # [n, n*]
        k_ss = _dense(self.covar_module(pts)).detach()
        mean = self.mean_module(x).detach() + k_star.t() @ alpha
        v = f.solve_lower(k_star)                                               # L^{-1} K_*  (one solve per test point)
        cov = k_ss + 1e-6 * torch.eye(x.shape[0], device=k_ss.device, dtype=dtype) - v.t() @ v
        return MultivariateNormal(mean, cov)


class ExactMarginalLogLikelihood(torch.nn.Module):
    """log p(target) / num_data under likelihood(output) (gpytorch.mlls.ExactMarginalLogLikelihood)."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output: MultivariateNormal, target):
        return self.likelihood(output).log_prob(target) / target.shape[0]


def train(model: ExactGPModel, train_x, train_y, training_iter=900, lr_scheduler_step=300, lr=1, verbose=True):
    """`examples/gpr_model.py:44-92`: Adam on all parameters (kernel hyper-parameters, mean, noise), StepLR(gamma=0.1)."""
    model.train()
    model.likelihood.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    scheduler = StepLR(optimizer, step_size=lr_scheduler_step, gamma=0.1)
    mll = ExactMarginalLogLikelihood(model.likelihood, model)
    losses = []
    for i in range(training_iter):
        optimizer.zero_grad()
        output = model(train_x)
        loss = -mll(output, train_y)
        loss.backward()
        optimizer.step()
        scheduler.step()
        losses.append(loss.item())
        if verbose and i % 100 == 99:
            measure = model.covar_module.measure
            print('Iter %d/%d - Loss: %.3f   lengthscale: %.3f variance: %.3f   noise: %.3f, mean: %3f' % (
                i + 1, training_iter, losses[-1], measure.lengthscale.item(), measure.variance.item(),
                model.likelihood.noise.item(), model.mean_module.constant.item()))
    return losses
