"""`eigenspace.basis` (TranslatedCharactersBasis, reference space.py:284-314) and EigenbasisKernel (spectral_kernel.py:57-75): golden
outputs of the real reference with its translations injected, and the identity sum_k b_k(x) conj(b_k(y)) = d chi(x y^{-1}) that
makes EigenbasisKernel equal EigenbasisSumKernel."""
import pytest
import torch

from conftest import load_golden, scaled_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["basis_so3", "basis_su2", "basis_su3"])
def test_translated_characters_basis_golden(name):
    from liestationarykernels_b200.spaces import SO, SU
    g = load_golden(name)
    space = (SO if g["group"] == "SO" else SU)(g["n"], order=g["order"])
    x, y = g["x"].cuda(), g["y"].cuda()
    torch.manual_seed(0)
    for eig, ref in zip(space.lb_eigenspaces, g["irreps"]):
        assert tuple(eig.index) == tuple(ref["index"]) and eig.dimension == ref["dimension"]
        basis = eig.basis
        assert basis.translations.shape == ref["translations"].shape
        basis.translations = ref["translations"].cuda()
        basis._factorise()
        # the Gram of d^2 random translates is ill-conditioned (cond ~ 1e3..1e6): compare at 1e-8 of scale
        assert scaled_err(torch.view_as_real(basis.coeffs.cpu()), torch.view_as_real(ref["coeffs"])) < 1e-8
        fx, fy = basis(x).cpu(), basis(y).cpu()
        assert fx.shape == ref["f_x"].shape
        assert scaled_err(torch.view_as_real(fx), torch.view_as_real(ref["f_x"])) < 1e-8
        assert scaled_err(torch.view_as_real(fy), torch.view_as_real(ref["f_y"])) < 1e-8
        # reproducing property: sum_k |b_k(x)|^2 = d chi(e) = d^2
        assert torch.allclose((fx.abs() ** 2).sum(dim=1), torch.full((x.shape[0],), float(eig.dimension) ** 2, dtype=torch.float64), rtol=1e-8)


@pytest.mark.parametrize("group,n,order", [("SO", 3, 3), ("SU", 2, 4), ("SU", 3, 3)])
def test_eigenbasis_kernel_equals_eigenbasis_sum_kernel(group, n, order):
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisKernel, EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(4)
    space = (SO if group == "SO" else SU)(n, order=order)
    meas = MaternSpectralMeasure(space.dim, 0.9, 2.5, 1.0)
    explicit = EigenbasisKernel(meas, space)
    summed = EigenbasisSumKernel(meas, space).eval()
    x, y = space.rand(11), space.rand(6)
    with torch.no_grad():
        a, b = explicit(x, y), summed(x, y)
        assert a.shape == (11, 6)
        assert (a - b).abs().max().item() < 1e-8
        kxx = explicit(x)
        assert torch.allclose(kxx, kxx.T, atol=1e-10) and torch.allclose(kxx.diagonal(), torch.ones(11, dtype=torch.float64, device="cuda"), atol=1e-8)


@pytest.mark.parametrize("cls,n,order", [("Sphere", 2, 5), ("Sphere", 3, 4), ("Sphere", 5, 3), ("ProjectiveSpace", 2, 3)])
def test_spherical_harmonics_basis_and_eigenbasis_kernel_on_the_sphere(cls, n, order):
    """`eigenspace.basis` on S^n / RP^n (NormalizedSphericalFunctions, sphere.py:126-141): the right number of functions, the
    addition theorem sum_k Y_k(x) Y_k(y) = Z_l(<x, y>) (the basis-independent statement; the reference's basis comes from the
    absent `spherical_harmonics` package and its own fundamental system), and EigenbasisKernel = EigenbasisSumKernel."""
    from liestationarykernels_b200 import spaces
    from liestationarykernels_b200.spectral_kernel import EigenbasisKernel, EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(8)
    space = getattr(spaces, cls)(n, order=order)
    state = torch.random.get_rng_state()
    x, y = space.rand(13), space.rand(9)
    for eig in space.lb_eigenspaces:
        yx, yy = eig.basis(x), eig.basis(y)
        assert yx.shape == (13, eig.dimension)
        want = eig.phase_function((x @ y.T).reshape(-1)).reshape(13, 9)
        assert (yx @ yy.T - want).abs().max().item() < 1e-9 * max(1.0, want.abs().max().item())
    meas = MaternSpectralMeasure(space.dim, 0.8, 2.5, 1.0)
    a = EigenbasisKernel(meas, space)(x, y)
    b = EigenbasisSumKernel(meas, space).eval()(x, y)
    assert (a - b).abs().max().item() < 1e-9
