"""GP regression step on the covariance: the blocked fp64 Cholesky / triangular solves behind the C ABI
(`lsk_potrf_upper`, `lsk_potrs_upper`) against LAPACK on the CPU, and the marginal likelihood, its gradients and the
posterior of `liestationarykernels_b200.gpr` against the oracle's restatement of the gpytorch path.

Tolerances (float64): factor and solves <= 1e-10 relative (matrices with condition number <= 1e5), log-likelihood and
gradients <= 1e-9 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _covariance(n, noise, seed=0):
    """a real covariance matrix of the hot path: SO(3) heat kernel at n Haar points + noise"""
    from liestationarykernels_b200.spaces import SO
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
    torch.manual_seed(seed)
    space = SO(3, order=20)
    kern = EigenbasisSumKernel(SqExpSpectralMeasure(3, 0.7), space).eval()
    with torch.no_grad():
        k = kern(space.rand(n))
    return k + noise * torch.eye(n, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("n", [1, 3, 64, 127, 128, 129, 255, 256, 257, 300, 385, 777, 1025])
def test_cholesky_matches_lapack(n):
    from liestationarykernels_b200.linalg import CholeskyFactor
    a = _covariance(n, 1e-2, seed=n)
    ref = torch.linalg.cholesky(a.cpu())
    f = CholeskyFactor(a)
    assert f.info == 0
    assert (f.L.cpu() - ref).abs().max().item() <= 1e-10 * ref.abs().max().item()
    assert torch.equal(f.U, f.L.mT)
    torch.manual_seed(1)
    b = torch.randn(n, dtype=torch.float64)
    x_ref = torch.cholesky_solve(b[:, None], ref)[:, 0]
    y_ref = torch.linalg.solve_triangular(ref, b[:, None], upper=False)[:, 0]
    z_ref = torch.linalg.solve_triangular(ref.mT, b[:, None], upper=True)[:, 0]
    for got, want in ((f.solve(b.cuda()), x_ref), (f.solve_lower(b.cuda()), y_ref), (f.solve_upper(b.cuda()), z_ref)):
        assert (got.cpu() - want).abs().max().item() <= 1e-10 * want.abs().max().item()
    assert abs(f.logdet().item() - 2 * torch.log(torch.diagonal(ref)).sum().item()) <= 1e-10 * max(1.0, n)
    inv_ref = torch.cholesky_inverse(ref)
    inv = f.inverse()
    assert torch.equal(inv, inv.mT)
    assert (inv.cpu() - inv_ref).abs().max().item() <= 1e-10 * inv_ref.abs().max().item()


def test_cholesky_multiple_right_hand_sides_and_strided_input():
    from liestationarykernels_b200.linalg import CholeskyFactor, cholesky
    a = _covariance(300, 1e-2)
    big = torch.zeros(300, 340, dtype=torch.float64, device="cuda")
    big[:, :300] = a
    view = big[:, :300]                              # row stride 340: factorised from a copy, the input is untouched
    f = CholeskyFactor(view)
    assert torch.equal(big[:, :300], a)
    b = torch.randn(300, 5, dtype=torch.float64, device="cuda")
    x = f.solve(b)
    assert ((a @ x - b).abs().max() / b.abs().max()).item() < 1e-10
    ref = torch.linalg.cholesky(a.cpu())
    for nrhs in (5, 64, 131, 300):                     # blocked many-right-hand-side path (lsk_trsm_lower) from 5 columns on
        bb = torch.randn(300, nrhs, dtype=torch.float64, device="cuda")
        want = torch.linalg.solve_triangular(ref, bb.cpu(), upper=False)
        got = f.solve_lower(bb)
        assert (got.cpu() - want).abs().max().item() <= 1e-10 * want.abs().max().item()
    L = cholesky(a)
    assert (L @ L.mT - a).abs().max().item() < 1e-12
    assert (cholesky(a, upper=True) - L.mT).abs().max().item() == 0.0


def test_cholesky_in_place_with_leading_dimension():
    """C ABI with lda > n: the factor lands in the upper triangle of the caller's buffer, columns beyond n are untouched"""
    import ctypes
    from liestationarykernels_b200 import _lib
    lib = _lib.load()
    n, lda = 333, 352
    a = _covariance(n, 1e-2)
    buf = torch.full((n, lda), 7.0, dtype=torch.float64, device="cuda")
    buf[:, :n] = a
    work = torch.empty(lib.lsk_potrf_work_doubles(n), dtype=torch.float64, device="cuda")
    info = torch.ones(1, dtype=torch.int32, device="cuda")
    assert lib.lsk_potrf_upper(_lib.ptr(buf), n, lda, _lib.ptr(work), _lib.ptr(info), _lib.stream_ptr()) == 0
    assert info.item() == 0
    ref = torch.linalg.cholesky(a.cpu())
    assert (torch.triu(buf[:, :n]).cpu() - ref.mT).abs().max().item() <= 1e-10 * ref.abs().max().item()
    assert torch.all(buf[:, n:] == 7.0)
    # argument errors
    assert lib.lsk_potrf_upper(_lib.ptr(buf), n, n - 1, _lib.ptr(work), _lib.ptr(info), _lib.stream_ptr()) == -2
    assert lib.lsk_potrf_upper(None, n, lda, _lib.ptr(work), _lib.ptr(info), _lib.stream_ptr()) == -1
    assert lib.lsk_potrs_upper(_lib.ptr(buf), n, lda, _lib.ptr(work), _lib.ptr(work), _lib.ptr(work), 0, _lib.stream_ptr()) == -1
    # inverse through the C ABI into a padded output
    iw = torch.empty(lib.lsk_potri_work_doubles(n), dtype=torch.float64, device="cuda")
    out = torch.full((n, lda), 3.0, dtype=torch.float64, device="cuda")
    assert lib.lsk_potri_upper(_lib.ptr(buf), n, lda, _lib.ptr(work), _lib.ptr(iw), _lib.ptr(out), lda, _lib.stream_ptr()) == 0
    inv_ref = torch.cholesky_inverse(ref)
    assert (out[:, :n].cpu() - inv_ref).abs().max().item() <= 1e-10 * inv_ref.abs().max().item()
    assert torch.all(out[:, n:] == 3.0)
    assert lib.lsk_potri_upper(_lib.ptr(buf), n, lda, _lib.ptr(work), _lib.ptr(iw), _lib.ptr(out), n - 1, _lib.stream_ptr()) == -2


def test_not_positive_definite_reports_lapack_info():
    from liestationarykernels_b200.linalg import CholeskyFactor, NotPositiveDefiniteError
    a = _covariance(300, 1e-2)
    a[200, 200] = -1.0
    assert CholeskyFactor(a, check=False).info == 201
    with pytest.raises(NotPositiveDefiniteError):
        CholeskyFactor(a)
    empty = CholeskyFactor(torch.zeros(0, 0, dtype=torch.float64, device="cuda"))
    assert empty.info == 0 and empty.L.shape == (0, 0)


def test_large_factorisation_residual():
    """size-independent property at a size LAPACK would take too long for: ||U^T U - A|| and A x = b residuals"""
    from liestationarykernels_b200.linalg import CholeskyFactor
    n = 5000
    a = _covariance(n, 1e-1)
    f = CholeskyFactor(a)
    u = f.U
    assert ((u.mT @ u - a).abs().max() / a.abs().max()).item() < 1e-13
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    assert ((a @ f.solve(b) - b).norm() / b.norm()).item() < 1e-12
    inv = f.inverse()
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert (a @ inv - eye).abs().max().item() < 1e-11
    bb = torch.randn(n, 777, dtype=torch.float64, device="cuda")
    v = f.solve_lower(bb)
    assert ((u.mT @ v - bb).abs().max() / bb.abs().max()).item() < 1e-12


def test_log_prob_and_gradients_match_lapack_autograd():
    from liestationarykernels_b200.gpr import MultivariateNormal
    import lsk_oracle as orc
    torch.manual_seed(3)
    n = 200
    a = _covariance(n, 5e-2).cpu()
    y = torch.randn(n, dtype=torch.float64)
    m = torch.full((n,), 0.3, dtype=torch.float64)
    a_ref, m_ref = a.clone().requires_grad_(True), m.clone().requires_grad_(True)
    want = orc.gp_log_marginal(a_ref, torch.tensor(-1e-6, dtype=torch.float64), m_ref, y) * n
    want.backward()
    a_g, m_g = a.cuda().requires_grad_(True), m.cuda().requires_grad_(True)
    got = MultivariateNormal(m_g, a_g).log_prob(y.cuda())
    got.backward()
    assert abs(got.item() - want.item()) <= 1e-10 * abs(want.item())
    # autograd through LAPACK's potrf gives the gradient of the lower triangle only; symmetrise both
    g_ref = 0.5 * (a_ref.grad + a_ref.grad.T)
    g_got = 0.5 * (a_g.grad + a_g.grad.mT).cpu()
    assert (g_got - g_ref).abs().max().item() <= 1e-9 * g_ref.abs().max().item()
    assert (m_g.grad.cpu() - m_ref.grad).abs().max().item() <= 1e-9 * m_ref.grad.abs().max().item()


def test_gp_training_step_matches_oracle():
    """one evaluation of the reference's training loss (examples/gpr_model.py:59-66) and its gradients with respect to
    lengthscale, variance, noise and mean: ours on the GPU against the oracle kernel + LAPACK differentiated on the CPU"""
    import lsk_oracle as orc
    from liestationarykernels_b200 import gpr
    from liestationarykernels_b200.spaces import SO
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(5)
    og = orc.Group("SO", 3, 10)
    n = 40
    x = og.rand(n)
    y = torch.randn(n, dtype=torch.float64)
    # ---- oracle
    om = orc.Measure("matern", og.dim, 0.8, 1.5, 1.7)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    raw_noise = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    const = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    emb = og.pairwise_embed(x, x)
    cov = torch.zeros(n, n, dtype=torch.float64)
    norm = torch.zeros((), dtype=torch.float64)
    for irrep in og.irreps:
        cov = cov + om(irrep[2]) * og.phase_function(irrep, emb).view(n, n).real
        norm = norm + om(irrep[2])[0] * irrep[1] ** 2
    k_ref = torch.abs(om.variance[0]) * cov / norm
    loss_ref = -orc.gp_log_marginal(k_ref, orc.gp_noise(raw_noise)[0], const.expand(n), y)
    loss_ref.backward()
    # ---- ours
    space = SO(3, order=10)
    meas = MaternSpectralMeasure(og.dim, 0.8, 1.5, 1.7)
    kern = EigenbasisSumKernel(meas, space)
    lik = gpr.GaussianLikelihood(device="cuda")
    tx = x.cuda().reshape(n, -1)
    model = gpr.ExactGPModel(tx, y.cuda(), lik, kern, space, point_shape=(3, 3))
    model.train()
    loss = -gpr.ExactMarginalLogLikelihood(lik, model)(model(tx), y.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-10 * abs(loss_ref.item())
    pairs = ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad),
             (lik.raw_noise.grad, raw_noise.grad), (model.mean_module.constant.grad, const.grad))
    for ours, ref in pairs:
        assert ours is not None
        assert abs(ours.item() - ref.item()) <= 1e-9 * max(1e-3, abs(ref.item())), (ours.item(), ref.item())


def test_gp_posterior_and_training_loop():
    """train() lowers the loss; the eval-mode posterior equals the oracle's LAPACK posterior for the trained parameters"""
    import lsk_oracle as orc
    from liestationarykernels_b200 import gpr
    from liestationarykernels_b200.spaces import SO
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    torch.manual_seed(7)
    space = SO(3, order=10)
    f = RandomPhaseApproximation(EigenbasisSumKernel(MaternSpectralMeasure(3, 1.0, 2.5, 4.0), space), phase_order=200)
    n_train, n_test = 200, 50
    train_x, test_x = space.rand(n_train), space.rand(n_test)
    with torch.no_grad():
        train_y, test_y = f(train_x), f(test_x)
    meas = MaternSpectralMeasure(3, 1.0, 1.5)
    kern = EigenbasisSumKernel(meas, space)
    lik = gpr.GaussianLikelihood(device="cuda")
    model = gpr.ExactGPModel(train_x.reshape(n_train, -1), train_y, lik, kern, space, point_shape=(3, 3))
    losses = gpr.train(model, train_x.reshape(n_train, -1), train_y, training_iter=30, lr_scheduler_step=15, lr=0.1, verbose=False)
    assert losses[-1] < losses[0]
    model.eval()
    kern.eval()
    with torch.no_grad():
        pred = model(test_x.reshape(n_test, -1))
        k = kern(train_x).cpu()
        ks = kern(train_x, test_x).cpu()
        kss = kern(test_x).cpu()
    mean_ref, cov_ref = orc.gp_posterior(k, ks, kss, lik.noise.detach().cpu()[0], model.mean_module.constant.detach().cpu(),
                                         train_y.cpu())
    assert (pred.mean.cpu() - mean_ref).abs().max().item() <= 1e-9 * max(1.0, mean_ref.abs().max().item())
    assert (pred.covariance_matrix.cpu() - cov_ref).abs().max().item() <= 1e-9 * max(1.0, cov_ref.abs().max().item())
    mse = ((pred.mean - test_y) ** 2).mean().item()
    assert mse < test_y.var(unbiased=False).item()          # the fitted GP beats the constant predictor
    test_ll = gpr.ExactMarginalLogLikelihood(lik, model)(pred, test_y).item()
    assert test_ll == test_ll


@pytest.mark.parametrize("kind", ["stiefel_random_phase", "hyperbolic_rff", "sphere"])
def test_gp_on_the_other_kernel_families(kind):
    """examples/gpr_experiments.py:31-73 picks the kernel class by space; the GP wrapper must accept each of them (lazy Gram
    results included); its posterior at the training inputs is checked against the closed form with a dense library solve"""
    from liestationarykernels_b200 import gpr
    from liestationarykernels_b200.spaces import HyperbolicSpace, Sphere, Stiefel
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomFourierFeatureKernel, RandomPhaseKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(9)
    if kind == "stiefel_random_phase":
        space = Stiefel(5, 3)
        kern = RandomPhaseKernel(MaternSpectralMeasure(space.dim, 1.0, 2.5), space, phase_order=64)
        shape = (5, 3)
    elif kind == "hyperbolic_rff":
        space = HyperbolicSpace(3, order=512)
        kern = RandomFourierFeatureKernel(MaternSpectralMeasure(space.dim, 1.0, 2.5), space)
        shape = None
    else:
        space = Sphere(3, order=12)
        kern = EigenbasisSumKernel(MaternSpectralMeasure(space.dim, 1.0, 2.5), space)
        shape = None
    n = 40
    x = space.rand(n).contiguous()
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    lik = gpr.GaussianLikelihood(device="cuda")
    flat = x.reshape(n, -1)
    model = gpr.ExactGPModel(flat, y, lik, kern, space, point_shape=shape)
    model.train()
    loss = -gpr.ExactMarginalLogLikelihood(lik, model)(model(flat), y)
    loss.backward()
    assert torch.isfinite(loss) and lik.raw_noise.grad is not None and model.mean_module.constant.grad is not None
    # lengthscale gradients flow through EigenbasisSumKernel, RandomPhaseKernel and the hyperbolic Fourier-feature kernel
    assert kern.measure.lengthscale.grad is not None and torch.isfinite(kern.measure.lengthscale.grad).all()
    with torch.no_grad():
        lik.raw_noise.fill_(-12.0)                      # noise -> 1e-4 + 6e-6
    model.eval()
    kern.eval()
    with torch.no_grad():
        pred = model(flat)
    k = kern(x)
    k = k.evaluate() if hasattr(k, "evaluate") else k
    # the posterior at the training inputs in closed form (dense solve by the library as the independent check)
    m = model.mean_module.constant.detach()
    s = k + (lik.noise.detach() + 1e-6) * torch.eye(n, dtype=torch.float64, device="cuda")
    want_mean = m + k @ torch.linalg.solve(s, y - m)
    want_cov = k + 1e-6 * torch.eye(n, dtype=torch.float64, device="cuda") - k @ torch.linalg.solve(s, k)
    scale = k.diagonal().mean().item()
    assert (pred.mean - want_mean).abs().max().item() < 1e-7 * max(1.0, y.abs().max().item())
    assert (pred.covariance_matrix - want_cov).abs().max().item() < 1e-7 * scale
    assert pred.variance.min().item() > -1e-9 * scale


def test_likelihood_and_mll_are_pure_and_idempotent():
    """`likelihood(output)` / `mll(output, y)` evaluated twice on the same `output` give the same value, and neither changes
    `output.covariance_matrix` (gpytorch / reference semantics: examples/gpr_model.py:59-66 may log the loss twice)."""
    from liestationarykernels_b200 import gpr
    from liestationarykernels_b200.spaces import SO
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(11)
    space = SO(3, order=10)
    kern = EigenbasisSumKernel(MaternSpectralMeasure(3, 0.8, 1.5, 1.7), space)
    n = 150
    x = space.rand(n)
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    lik = gpr.GaussianLikelihood(device="cuda")
    model = gpr.ExactGPModel(x.reshape(n, -1), y, lik, kern, space, point_shape=(3, 3))
    model.train()
    mll = gpr.ExactMarginalLogLikelihood(lik, model)
    output = model(x.reshape(n, -1))
    before = output.covariance_matrix.detach().clone()
    a = mll(output, y)
    b = mll(output, y)
    noisy1 = lik(output).covariance_matrix.detach()
    noisy2 = lik(output).covariance_matrix.detach()
    assert a.item() == b.item()
    assert torch.equal(output.covariance_matrix.detach(), before)
    assert torch.equal(noisy1, noisy2)
    assert torch.allclose(noisy1.diagonal() - before.diagonal(), lik.noise.detach().expand(n), rtol=0, atol=1e-14)
    (a + b).backward()                                    # both graphs are intact
    assert torch.isfinite(kern.measure.lengthscale.grad).all() and torch.isfinite(lik.raw_noise.grad).all()


@pytest.mark.parametrize("n,k", [(300, 24), (1000, 129), (517, 64)])
def test_solve_with_many_right_hand_sides(n, k):
    """A^{-1} B for a block of right-hand sides (dense-inverse + Gram path from max(8, n / 256) columns on) against the per-column sweeps and
    against the residual"""
    from liestationarykernels_b200.linalg import CholeskyFactor
    torch.manual_seed(n + k)
    a = _covariance(n, 5e-2)
    f = CholeskyFactor(a)
    b = torch.randn(n, k, dtype=torch.float64, device="cuda")
    x = f.solve(b)
    assert x.shape == (n, k)
    assert ((a @ x - b).abs().max() / b.abs().max()).item() < 1e-10
    cols = torch.stack([f.solve(b[:, c].contiguous()) for c in range(0, k, max(k // 5, 1))], dim=1)
    assert (x[:, ::max(k // 5, 1)] - cols).abs().max().item() < 1e-10 * cols.abs().max().item()


@pytest.mark.parametrize("name", ["gp_so3_matern", "gp_su3_heat"])
def test_gp_step_against_committed_golden(name):
    """One training loss with all four gradients and the eval-mode posterior against committed fixtures: covariance and its
    hyper-parameter derivatives from the REAL reference kernel (torch autograd through its forward), marginal likelihood and
    posterior from the LAPACK restatement of gpytorch's dense path (`oracle/make_golden.py`, section GP regression step)."""
    from conftest import load_golden
    from liestationarykernels_b200 import gpr
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    g = load_golden(name)
    ms = g["measure"]
    space = (SO if g["group"] == "SO" else SU)(g["n"], order=g["order"])
    meas = MaternSpectralMeasure(ms["dim"], ms["lengthscale"], ms["nu"], ms["variance"]) if ms["kind"] == "matern" else \
        SqExpSpectralMeasure(ms["dim"], ms["lengthscale"], ms["variance"])
    kern = EigenbasisSumKernel(meas, space)
    lik = gpr.GaussianLikelihood(device="cuda")
    x, xs, targets = g["x"].cuda(), g["xs"].cuda(), g["targets"].cuda()
    n = x.shape[0]
    flat = (torch.view_as_real(x) if torch.is_complex(x) else x).reshape(n, -1)
    shape = (g["n"], g["n"], 2) if torch.is_complex(x) else (g["n"], g["n"])
    model = gpr.ExactGPModel(flat, targets, lik, kern, space, point_shape=shape)
    if torch.is_complex(x):
        model._points = lambda t: torch.view_as_complex(t.view(t.shape[0], g["n"], g["n"], 2).contiguous())
    with torch.no_grad():
        lik.raw_noise.copy_(g["raw_noise"].cuda())
        model.mean_module.constant.copy_(g["constant"].cuda())
    model.train()
    out = model(flat)
    assert (out._base.detach().cpu() - g["k_train"].detach()).abs().max().item() < 1e-10
    loss = -gpr.ExactMarginalLogLikelihood(lik, model)(out, targets)
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) <= 1e-10 * abs(g["loss"].item())
    for ours, ref in ((meas.lengthscale.grad, g["grad_lengthscale"]), (meas.variance.grad, g["grad_variance"]),
                      (lik.raw_noise.grad, g["grad_raw_noise"]), (model.mean_module.constant.grad, g["grad_constant"])):
        assert abs(ours.item() - ref.item()) <= 1e-8 * max(1e-3, abs(ref.item())), (ours.item(), ref.item())
    model.eval()
    kern.eval()
    kern.normalizer = g["eval_normalizer"].detach().cuda()
    xs_flat = (torch.view_as_real(xs) if torch.is_complex(xs) else xs).reshape(xs.shape[0], -1)
    with torch.no_grad():
        pred = model(xs_flat)
    assert (pred.mean.cpu() - g["posterior_mean"]).abs().max().item() <= 1e-9 * max(1.0, g["posterior_mean"].abs().max().item())
    assert (pred.covariance_matrix.cpu() - g["posterior_cov"]).abs().max().item() <= 1e-9


def test_pipelined_solve_with_fewer_ctas_than_blocks():
    """`lsk_potrs_upper` hands the 128-row blocks to its CTAs by a ticket, one block per round: with three CTAs for eleven blocks
    (LSK_SWEEP_CTAS, read once per process, hence the child process) every CTA goes through several rounds and the result is
    bit-for-bit the one of the default launch."""
    import os, subprocess, sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from liestationarykernels_b200.linalg import CholeskyFactor
torch.manual_seed(3)
n = 1337
g = torch.randn(n, n + 9, dtype=torch.float64, device="cuda")
a = g @ g.T / n + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
b = torch.randn(n, dtype=torch.float64, device="cuda")
f = CholeskyFactor(a)
x = f.solve(b)
assert ((a @ x - b).abs().max() / b.abs().max()).item() < 1e-11
print(x.cpu().numpy().tobytes().hex()[:0] + str(float(x.double().sum().item()).hex()) + " " + str(float((x * x).sum().item()).hex()))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for ctas in ("3", "0"):
        env = dict(os.environ, LSK_SWEEP_CTAS=ctas)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
