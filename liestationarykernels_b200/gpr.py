"""Exact GP regression on top of the covariance kernels — the caller of the hot path in the reference
(`examples/gpr_model.py:12-92`, `examples/gpr_experiments.py:75-87`), without gpytorch (absent from the image; SURVEY.md
8f row f2).  Same model and training loop: constant mean, Gaussian likelihood with gpytorch's default noise
parametrisation (`softplus(raw_noise) + 1e-4`), `covar = kernel(x) + 1e-6 I`, exact marginal log-likelihood divided by
the number of targets, Adam + StepLR.

The marginal likelihood and the posterior run on this package's fp64 Cholesky (`linalg.CholeskyFactor` ->
`lsk_potrf_upper` / `lsk_potrs_upper` / `lsk_potri_upper`, csrc/lsk_cholesky.cu).  Hyper-parameter gradients flow through
`EigenbasisSumKernel`'s fused backward (spectral_kernel._WeightedCharacterSum); the gradient of the log-determinant needs
the dense inverse (`CholeskyFactor.inverse`).

One deliberate difference from gpytorch: inside an autograd graph the diagonal shifts (`+ 1e-6 I`, `+ noise I`) are applied
IN PLACE to the covariance `ExactGPModel.forward` just produced (an N x N identity and two extra passes per shift otherwise),
so `output.covariance_matrix` of a training-mode `model(x)` holds the shifted matrix after `mll(output, y)` was evaluated.
Matrices the caller built (a `MultivariateNormal` of their own) and tensors outside a graph are never modified."""
from __future__ import annotations

import math

import torch
from torch.optim.lr_scheduler import StepLR

from .linalg import CholeskyFactor

dtype = torch.float64


def _dense(k):
    return k.evaluate() if hasattr(k, "evaluate") else k


class _GaussianLogProb(torch.autograd.Function):
    """log N(resid; 0, covar) for a dense positive-definite covar, with gradients
       d/dcovar = (alpha alpha^T - covar^{-1}) / 2,   d/dresid = -alpha,   alpha = covar^{-1} resid."""

    @staticmethod
    def forward(ctx, covar, resid):
        f = CholeskyFactor(covar, overwrite=False)
        alpha = f.solve(resid)
        n = resid.shape[0]
        ctx.factor, ctx.alpha = f, alpha
        return -0.5 * (torch.dot(resid, alpha) + f.logdet() + n * math.log(2.0 * math.pi))

    @staticmethod
    def backward(ctx, g):
        f, alpha = ctx.factor, ctx.alpha
        grad_covar = grad_resid = None
        if ctx.needs_input_grad[0]:
            gs = float(g)
            grad_covar = f.inverse().addr_(alpha, alpha, beta=-0.5 * gs, alpha=0.5 * gs)     # one pass over the N x N inverse
        if ctx.needs_input_grad[1]:
            grad_resid = -g * alpha
        return grad_covar, grad_resid


class _AddToDiagonal(torch.autograd.Function):
    """covar + shift * I, written into covar itself when this module owns the buffer (the kernel's fresh output): the
    reference's `K + 1e-6 * eye` and the likelihood's `+ noise * eye` each cost an N x N identity plus two passes over the
    matrix otherwise.  Gradients: identity for covar, trace for shift."""

    @staticmethod
    def forward(ctx, covar, shift):
        ctx.mark_dirty(covar)
        covar.diagonal().add_(shift.reshape(()) if isinstance(shift, torch.Tensor) else shift)
        return covar

    @staticmethod
    def backward(ctx, grad):
        g_shift = torch.diagonal(grad).sum().reshape(1) if ctx.needs_input_grad[1] else None
        return grad, g_shift


def _add_to_diagonal(covar, shift, owned):
    """covar + shift * I; in place only when `owned` says the buffer was produced inside this module (the kernel's fresh
    output inside ExactGPModel.forward) and it is part of an autograd graph; any other tensor is copied first."""
    shift_t = shift if isinstance(shift, torch.Tensor) else torch.tensor([shift], dtype=dtype, device=covar.device)
    if not owned or covar.is_leaf or covar.stride(1) != 1:
        covar = covar.clone()
    return _AddToDiagonal.apply(covar, shift_t)


class MultivariateNormal:
    """The slice of `gpytorch.distributions.MultivariateNormal` the reference's scripts use."""

    def __init__(self, mean, covariance_matrix, owns_covariance=False):
        self.mean = mean
        self.covariance_matrix = covariance_matrix
        self.owns_covariance = owns_covariance          # the matrix was produced by this module: later shifts may reuse it

    loc = property(lambda self: self.mean)

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix)

    @property
    def stddev(self):
        return self.variance.clamp_min(0).sqrt()

    def log_prob(self, value):
        return _GaussianLogProb.apply(self.covariance_matrix, value - self.mean)


class GaussianLikelihood(torch.nn.Module):
    """Homoskedastic noise, gpytorch's default parametrisation: noise = softplus(raw_noise) + 1e-4, raw_noise = 0."""

    def __init__(self, device=None):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1, dtype=dtype, device=device))

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.raw_noise) + 1e-4

    def forward(self, dist: MultivariateNormal) -> MultivariateNormal:
        owned = getattr(dist, "owns_covariance", False)
        return MultivariateNormal(dist.mean, _add_to_diagonal(dist.covariance_matrix, self.noise, owned), owns_covariance=owned)


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
        covar_x = _add_to_diagonal(_dense(self.covar_module(self._points(x))), 1e-6, owned=True)
        return MultivariateNormal(mean_x, covar_x, owns_covariance=True)

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
