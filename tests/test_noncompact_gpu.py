"""GPU parity for H^n / SPD(n): spherical-function features, RandomFourierFeatureKernel, RandomFourierApproximation,
against golden outputs of the real reference with injected spectral samples, shifts and weights."""
import pytest
import torch

from conftest import cov_err, load_golden, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-10

CASES = ["rff_h3_matern", "rff_h3_heat", "rff_h2_heat", "rff_h5_matern", "rff_spd3_matern", "rff_spd3_heat",
         "rff_spd2_heat", "rff_spd4_heat"]


def _measure(spec):
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    if spec["kind"] == "matern":
        return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], spec.get("variance", 1.0))
    return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], spec.get("variance", 1.0))


@pytest.mark.parametrize("name", CASES)
def test_noncompact_golden(name):
    from liestationarykernels_b200 import spaces
    from liestationarykernels_b200.spaces.noncompact import _SphericalFunctions
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel, RandomSpectralKernel
    from liestationarykernels_b200.prior_approximation import RandomFourierApproximation
    g = load_golden(name)
    space = getattr(spaces, g["cls"])(g["n"], order=g["order"])
    meas = _measure(g["measure"])
    torch.manual_seed(0)
    kern = RandomFourierFeatureKernel(meas, space)
    kern.eval()
    assert space.lb_eigenspaces.exp.lmd.shape == g["lmd"].shape
    assert space.lb_eigenspaces.exp.shift.shape == g["shift"].shape
    # inject the reference's spectral samples and shifts; the c-function is recomputed by our code
    lmd, shift = g["lmd"].cuda(), g["shift"].cuda()
    coeff = space.c_function(lmd) if g["cls"] == "HyperbolicSpace" else space.c_function_tanh(lmd)
    assert scaled_err(coeff.cpu(), g["coeff"]) < 1e-13
    space.lb_eigenspaces = _SphericalFunctions(lmd, shift, space, coeff)
    x, y = g["x"].cuda(), g["y"].cuda()
    with torch.no_grad():
        feats = space.lb_eigenspaces(space.to_group(x)).cpu()
        assert feats.dtype == torch.complex128 and feats.shape == g["features"].shape
        assert scaled_err(feats, g["features"]) < TOL
        kern.compute_normalizer()
        assert abs(float(kern.normalizer) / float(g["normalizer"]) - 1) < 1e-12
        gram = kern(x, y).evaluate().cpu()
        assert cov_err(gram, g["gram"]) < TOL
        sampler = RandomFourierApproximation(kern)
        sampler.weights_real, sampler.weights_imag = g["w_re"].cuda(), g["w_im"].cuda()
        assert scaled_err(sampler(x).cpu(), g["sample"]) < TOL
        assert cov_err(sampler._cov(x, y).cpu(), g["cov"]) < TOL


def test_spd_to_group_and_inverse():
    from liestationarykernels_b200.spaces import SymmetricPositiveDefiniteMatrices
    space = SymmetricPositiveDefiniteMatrices(3, order=8)
    torch.manual_seed(0)
    x = space.rand(200)
    u = space.to_group(x)
    # Wishart-type samples can be badly conditioned; both factorizations are backward stable, so compare residuals
    assert torch.allclose(u.transpose(-1, -2) @ u, x, rtol=1e-13, atol=1e-13)
    assert torch.equal(u, torch.triu(u))
    ref = torch.linalg.cholesky(x, upper=True)
    assert ((u - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2))).median().item() < 1e-14
    eye = torch.eye(3, dtype=torch.float64, device='cuda')
    assert ((space.inv(x) @ x - eye).abs().amax(dim=(1, 2)) / torch.linalg.cond(x)).max().item() < 1e-14


def test_hyperbolic_rff_and_pairwise_kernels_agree():
    """reference tests/hyperbolic.py:27-38 (Monte-Carlo, atol 5e-2): feature-map kernel vs pairwise spectral kernel"""
    from liestationarykernels_b200.spaces import HyperbolicSpace
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel, RandomSpectralKernel
    from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
    torch.manual_seed(0)
    space = HyperbolicSpace(3, order=400000)
    meas = SqExpSpectralMeasure(3, 1.0)
    k1 = RandomFourierFeatureKernel(meas, space)
    k1.eval()
    k2 = RandomSpectralKernel(meas, space)
    k2.eval()
    x = space.rand(10) * 0.6
    with torch.no_grad():
        a, b = k1(x, x).evaluate(), k2(x, x)
    assert torch.allclose(torch.diagonal(b), torch.ones(10, dtype=torch.float64, device="cuda"), atol=1e-9)
    assert (a - b).abs().max().item() < 0.1, (a - b).abs().max().item()


@pytest.mark.parametrize("kind,n,m", [("matern", 3, 64), ("heat", 2, 48), ("matern", 4, 40)])
def test_hyperbolic_fourier_kernel_hyperparameter_gradients(kind, n, m):
    """RandomFourierFeatureKernel on H^n in training mode (what `examples/hyperbolic_GPR.ipynb` / gpr_experiments.py train):
    value, lengthscale and variance gradients against torch autograd through the oracle port — the spectral samples scale
    with the lengthscale (hyperbolic.py:43-52), the c-function and the normaliser move with them."""
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import HyperbolicSpace
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    torch.manual_seed(23)
    oh = orc.Hyperbolic(n, order=m)
    om = orc.Measure(kind, n, 0.8, 1.5 if kind == "matern" else None, 1.6)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    nx, ny = 33, 19
    x, y = oh.rand(nx), oh.rand(ny)
    lmd, shift = oh.generate(om)
    norm = orc.rff_normalizer(oh, lmd, shift)
    k_ref = orc.rff_kernel(oh, om, lmd, shift, x, y, norm)
    g_out = torch.randn(nx, ny, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    space = HyperbolicSpace(n, order=m)
    space.normalized_lmd, space.shift = oh.normalized_lmd.cuda(), oh.shift.cuda()      # same-tensor rule
    meas = MaternSpectralMeasure(n, 0.8, 1.5, 1.6) if kind == "matern" else SqExpSpectralMeasure(n, 0.8, 1.6)
    kern = RandomFourierFeatureKernel(meas, space)
    kern.train()
    k = kern(x.cuda(), y.cuda()).evaluate()
    assert cov_err(k.detach().cpu(), k_ref.detach()) < 1e-10
    assert abs(float(kern.normalizer) - norm.item()) < 1e-10 * abs(norm.item())
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-8 * max(1.0, abs(ref.item())), (ours.item(), ref.item())
    # symmetric call and eval mode
    kern.eval()
    kxx = kern(x.cuda()).evaluate()
    with torch.no_grad():
        plain = kern(x.cuda()).evaluate()
    assert kxx.requires_grad and scaled_err(kxx.detach().cpu(), plain.cpu()) < 1e-13


@pytest.mark.parametrize("kind,n,m,training", [("matern", 3, 40, True), ("heat", 2, 32, True), ("matern", 2, 24, False)])
def test_spd_fourier_kernel_hyperparameter_gradients(kind, n, m, training):
    """RandomFourierFeatureKernel on SPD(n): value and lengthscale / variance gradients against torch autograd through the
    oracle port.  SPD re-draws shifts and spectra on every training-mode forward (spd.py:28): the oracle's draws are injected
    through `space._draw` (same-tensor rule).  Eval mode: stored features and normaliser, gradient through the scale only."""
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import SymmetricPositiveDefiniteMatrices
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    torch.manual_seed(29)
    os_ = orc.SPD(n, order=m)
    om = orc.Measure(kind, os_.dim, 0.9, 1.5 if kind == "matern" else None, 1.4)
    nx, ny = 21, 17
    x, y = os_.rand(nx), os_.rand(ny)
    # raw draws in the reference's order, then lmd as a differentiable function of the lengthscale
    shift = os_.rand_phase(m)
    goe = orc.goe_spectrum(m, n)
    chi2 = torch.distributions.chi2.Chi2(2 * om.nu[0]).rsample((m,)) if kind == "matern" else None
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    if kind == "matern":
        scale = 1 / torch.sqrt((n ** 3 - n) / 48 + 2 * om.nu[0] / om.lengthscale[0])
        lmd = goe / torch.sqrt(chi2)[:, None] / scale
    else:
        lmd = goe / om.lengthscale
    norm = orc.rff_normalizer(os_, lmd, shift)
    if not training:
        norm = norm.detach()
    k_ref = orc.rff_kernel(os_, om, lmd, shift, x, y, norm)
    g_out = torch.randn(nx, ny, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    space = SymmetricPositiveDefiniteMatrices(n, order=m)
    space._draw = lambda measure: (shift.cuda(), goe.cuda(), None if chi2 is None else chi2.cuda())
    meas = MaternSpectralMeasure(os_.dim, 0.9, 1.5, 1.4) if kind == "matern" else SqExpSpectralMeasure(os_.dim, 0.9, 1.4)
    kern = RandomFourierFeatureKernel(meas, space)
    kern.train(training)
    k = kern(x.cuda(), y.cuda()).evaluate()
    assert cov_err(k.detach().cpu(), k_ref.detach()) < 1e-10
    assert abs(float(kern.normalizer) - norm.item()) < 1e-10 * abs(norm.item())
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-8 * max(1.0, abs(ref.item())), (ours.item(), ref.item())


@pytest.mark.parametrize("kind,n,m", [("matern", 3, 48), ("heat", 2, 40)])
def test_random_spectral_kernel_hyperparameter_gradients(kind, n, m):
    """RandomSpectralKernel (the pairwise form, spectral_kernel.py:130-161) is plain torch in the reference and therefore
    differentiable in lengthscale and variance; ours: value and both gradients against torch autograd through the oracle port."""
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import HyperbolicSpace
    from liestationarykernels_b200.spectral_kernel import RandomSpectralKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    torch.manual_seed(41)
    oh = orc.Hyperbolic(n, order=m)
    om = orc.Measure(kind, n, 0.8, 1.5 if kind == "matern" else None, 1.6)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    nx, ny = 13, 9
    x, y = oh.rand(nx) * 0.7, oh.rand(ny) * 0.7
    lmd, shift = oh.generate(om)
    norm = (oh.c_function(lmd) ** 2).sum()                   # the kernel at coinciding points: sum_k |F_k(id)|^2
    k_ref = orc.random_spectral_kernel(oh, om, lmd, shift, x, y, norm)
    g_out = torch.randn(nx, ny, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    space = HyperbolicSpace(n, order=m)
    space.normalized_lmd, space.shift = oh.normalized_lmd.cuda(), oh.shift.cuda()      # same-tensor rule
    meas = MaternSpectralMeasure(n, 0.8, 1.5, 1.6) if kind == "matern" else SqExpSpectralMeasure(n, 0.8, 1.6)
    kern = RandomSpectralKernel(meas, space)
    kern.train()
    k = kern(x.cuda(), y.cuda())
    assert k.shape == (nx, ny) and scaled_err(k.detach().cpu(), k_ref.detach()) < 1e-10
    assert abs(float(kern.normalizer) - norm.item()) < 1e-10 * abs(norm.item())
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-8 * max(1.0, abs(ref.item())), (ours.item(), ref.item())
    kern.eval()
    with torch.no_grad():
        plain = kern(x.cuda(), y.cuda())
    assert scaled_err(plain.cpu(), k_ref.detach()) < 1e-10


def test_random_spectral_kernel_on_spd_is_differentiable():
    from liestationarykernels_b200.spaces import SymmetricPositiveDefiniteMatrices
    from liestationarykernels_b200.spectral_kernel import RandomSpectralKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(43)
    space = SymmetricPositiveDefiniteMatrices(2, order=32)
    meas = MaternSpectralMeasure(space.dim, 0.9, 1.5, 1.3)
    kern = RandomSpectralKernel(meas, space)
    kern.eval()
    x = space.rand(7)
    k = kern(x, x)
    assert k.requires_grad and torch.isfinite(k).all()
    with torch.no_grad():
        plain = kern(x, x)
    assert scaled_err(k.detach(), plain) < 1e-12
    k.sum().backward()
    assert torch.isfinite(meas.lengthscale.grad).all()
    # variance enters linearly: dK/dvar = K / var
    assert abs(meas.variance.grad.item() - (plain.sum() / 1.3).item()) < 1e-9 * abs(plain.sum().item())
