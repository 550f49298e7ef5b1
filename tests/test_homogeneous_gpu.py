"""GPU parity for Stiefel / Grassmannian / oriented Grassmannian: lift, feature map, Gram, EigenbasisSumKernel and the
sampler, against golden outputs of the real reference with injected phases / weights / subgroup samples."""
import pytest
import torch

from conftest import cov_err, load_golden, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-10

CASES = ["homog_stiefel53", "homog_stiefel52_heat", "homog_grassmannian52", "homog_orgrassmannian52", "homog_stiefel32",
         "homog_stiefel42", "homog_grassmannian42", "homog_orgrassmannian42", "homog_grassmannian31"]


def _measure(spec):
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    if spec["kind"] == "matern":
        return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], spec.get("variance", 1.0))
    return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], spec.get("variance", 1.0))


@pytest.mark.parametrize("name", CASES)
def test_homogeneous_golden(name):
    from liestationarykernels_b200 import spaces
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomPhaseKernel
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    g = load_golden(name)
    space = getattr(spaces, g["cls"])(*g["nm"], **g["kw"])
    assert space.h_samples.shape == g["h_samples"].shape
    space.h_samples = g["h_samples"].cuda()
    meas = _measure(g["measure"])
    x, y = g["x"].cuda(), g["y"].cuda()
    P = g["phases"].shape[0]
    with torch.no_grad():
        lift = space.M_to_G(x).cpu()
        assert (lift - g["lift_x"]).abs().max().item() < 1e-13
        kern = RandomPhaseKernel(meas, space, phase_order=P)
        kern.eval()
        kern.phases = g["phases"].cuda()
        emb = kern.make_embedding(x).cpu()
        assert scaled_err(emb, g["embedding"]) < TOL
        raw = kern(x, y, normalize=False).evaluate().cpu()
        assert cov_err(raw, g["gram_raw"]) < TOL
        kern.normalizer = g["normalizer"].cuda()
        gram = kern(x, y).evaluate().cpu()
        assert cov_err(gram, g["gram"]) < TOL
        esum = EigenbasisSumKernel(meas, space)
        esum.eval()
        esum.normalizer = g["es_normalizer"].cuda()
        kes = esum(x, y).cpu()
        assert cov_err(kes, g["k_es"]) < TOL
        sampler = RandomPhaseApproximation(esum, phase_order=P)
        sampler.phases = g["s_phases"].cuda()
        sampler.weights = g["s_weights"].cuda()
        assert scaled_err(sampler.make_embedding(x).cpu(), g["s_embedding"]) < TOL
        assert scaled_err(sampler(x).cpu(), g["sample"]) < TOL


def test_points_have_unit_columns_and_lift_is_special_orthogonal():
    """reference tests/homogeneous_spaces.py:59-61 + the asserts of grassmannian.py:84-85"""
    from liestationarykernels_b200.spaces import Stiefel
    space = Stiefel(5, 3, order=10, average_order=10)
    torch.manual_seed(1)
    x = space.rand(500)
    assert torch.allclose(torch.linalg.norm(x, dim=1), torch.ones(500, 3, dtype=torch.float64, device="cuda"))
    g = space.M_to_G(x)
    eye = torch.eye(5, dtype=torch.float64, device="cuda")
    assert torch.allclose(g.transpose(-1, -2) @ g, eye.expand(500, 5, 5), atol=1e-13)
    assert torch.allclose(torch.det(g), torch.ones(500, dtype=torch.float64, device="cuda"), atol=1e-12)
    assert torch.allclose(g[:, :, :3], x, atol=1e-13)


def test_prior_covariance_matches_kernel_statistically():
    """reference tests/homogeneous_spaces.py:71-74: sampler._cov ~ EigenbasisSumKernel.  Monte-Carlo statement (the
    reference uses atol 5e-2 with average_order=1000); with the cheaper settings here the bound is 0.12."""
    from liestationarykernels_b200.spaces import Grassmannian
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    torch.manual_seed(0)
    space = Grassmannian(5, 2, order=10, average_order=400)
    kern = EigenbasisSumKernel(SqExpSpectralMeasure(6, 1.0), space)
    kern.eval()
    sampler = RandomPhaseApproximation(kern, phase_order=2000)
    x = space.rand(12)
    with torch.no_grad():
        assert torch.allclose(sampler._cov(x, x), kern(x, x), atol=0.12)


@pytest.mark.parametrize("kind,n,m,order,avg", [
    ("stiefel", 5, 2, 10, 12), ("stiefel", 5, 4, 10, 6), ("stiefel", 5, 3, 20, 10), ("grassmannian", 5, 2, 10, 12),
    ("oriented_grassmannian", 5, 3, 10, 9), ("grassmannian", 5, 3, 20, 8), ("stiefel", 5, 1, 10, 8), ("stiefel", 5, 3, 26, 6)])
def test_so5_quotients_against_oracle(kind, n, m, order, avg):
    """SO(5)/H through every instantiation of `homog5_stair_kernel` (staircase of the first 10 / 20 irreps x subgroup samples
    of the form blockdiag(I_3, rotation) [V(5,3), V(5,4): `homog5_rot_kernel`], blockdiag(I_2, R) [V(5,2)] or general [Grassmannians, V(5,1)]) and the
    generic fallback (26 irreps): EigenbasisSumKernel and the per-irrep feature map against the oracle on the same points and
    the same subgroup samples, ragged sizes."""
    import lsk_oracle as orc
    from liestationarykernels_b200 import spaces
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomPhaseKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(31)
    oh = orc.Homogeneous(kind, n, m, order=order, average_order=avg)
    nx, ny, P = 37, 29, 12
    x, y, phases = oh.rand(nx), oh.rand(ny), oh.rand(P)
    om = orc.Measure("matern", oh.dim, 0.9, 2.5, 1.3)
    cls = {"stiefel": spaces.Stiefel, "grassmannian": spaces.Grassmannian, "oriented_grassmannian": spaces.OrientedGrassmannian}[kind]
    space = cls(n, m, order=order, average_order=avg)
    space.h_samples = oh.h_samples.cuda()
    # structure code of the subgroup samples: blockdiag(I_2, R) -> 2; blockdiag(I_3, plane rotation) -> 11 (V(5,3), V(5,4))
    expect_block = {("stiefel", 2): 2, ("stiefel", 3): 11, ("stiefel", 4): 11}.get((kind, m), 0)
    assert space._identity_block(space.h_samples) == expect_block
    meas = MaternSpectralMeasure(oh.dim, 0.9, 2.5, 1.3)
    with torch.no_grad():
        kern = EigenbasisSumKernel(meas, space).eval()
        ref = orc.eigensum_unnormalised(oh, om, x, y)
        got = kern(x.cuda(), y.cuda(), normalize=False).cpu()
        assert cov_err(got, ref) < 1e-10
        rp = RandomPhaseKernel(meas, space, phase_order=P).eval()
        rp.phases = phases.cuda()
        emb_ref = orc.random_phase_embedding(oh, om, phases, x, torch.sqrt)
        emb = rp.make_embedding(x.cuda()).cpu()
        assert scaled_err(emb, emb_ref) < 1e-10


@pytest.mark.parametrize("order,avg,P,N", [(10, 30, 64, 100), (20, 7, 36, 77), (10, 200, 8, 9)])
def test_fused_stiefel_sampler_equals_feature_map_times_weights(order, avg, P, N):
    """`lsk_homog_sample` (V(5,3): the feature kernel reduces over the phases itself) against the two-pass path (feature map written
    panel-major, then `lsk_panel_gemv`) and against the block-structured kernel for subgroup samples that are NOT plane rotations
    (structure code 3 instead of 11): same numbers to rounding; ragged sizes, more subgroup samples than one staging pass holds."""
    from liestationarykernels_b200 import ops
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    from liestationarykernels_b200.spaces import Stiefel
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(order + avg)
    space = Stiefel(5, 3, order=order, average_order=avg)
    assert space._identity_block(space.h_samples) == 11
    kern = EigenbasisSumKernel(MaternSpectralMeasure(9, 0.8, 2.5, 1.3), space).eval()
    sampler = RandomPhaseApproximation(kern, phase_order=P)
    x = space.rand(N)
    with torch.no_grad():
        fused = sampler(x)
        feats = space._phase_features(x, sampler.phases, sampler._row_scales(), panel=True)
        two_pass = ops.panel_matvec(feats, sampler.weights)
        dense = sampler.make_embedding(x)
        assert scaled_err(fused, two_pass) < 1e-13
        assert scaled_err(fused, dense @ sampler.weights) < 1e-13
        # the same samples with a structure the rotation kernel must refuse: scale R by 1 + 1e-9 (no longer a rotation)
        h2 = space.h_samples.clone()
        h2[:, 3:, 3:] *= 1.0 + 1e-9
        space.h_samples = h2
        assert space._identity_block(space.h_samples) == 3
        general = sampler(x)
        assert scaled_err(general, fused) < 1e-7 and not torch.equal(general, fused)
