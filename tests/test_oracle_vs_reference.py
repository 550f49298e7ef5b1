"""The oracle port (`oracle/lsk_oracle.py`) against the LIVE reference, imported in this container through
`oracle/ref_shim.py`, on fresh seeds that no golden file holds.  Skipped where `/root/reference` is absent (the GPU box):
there the committed goldens (`tests/test_oracle_golden.py`) carry the same pin.

Same-tensor rule: points / phases / weights / subgroup samples / spectral samples are drawn once (by the oracle, CPU, seeded)
and injected into the reference objects, which are plain-attribute classes."""
import warnings

import pytest
import torch

import ref_shim
import lsk_oracle as orc

from conftest import cov_err, rel_err, scaled_err

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
TOL = 1e-11


@pytest.fixture(scope="module")
def ref():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_shim.import_reference()
        import lie_stationary_kernels.spaces as spaces
        import lie_stationary_kernels.spectral_kernel as sk
        import lie_stationary_kernels.spectral_measure as sm
        import lie_stationary_kernels.prior_approximation as pa
    return dict(spaces=spaces, sk=sk, sm=sm, pa=pa)


def _ref_measure(ref, kind, dim, ls, nu, var):
    if kind == "matern":
        return ref["sm"].MaternSpectralMeasure(dim, ls, nu, var)
    return ref["sm"].SqExpSpectralMeasure(dim, ls, var)


@pytest.mark.parametrize("family,n,order,kind,ls,nu,var,nx,ny", [
    ("SO", 3, 30, "matern", 0.7, 1.5, 2.0, 23, 17),
    ("SO", 5, 20, "heat", 0.9, None, 1.0, 9, 11),
    ("SO", 4, 12, "matern", 1.1, 2.5, 0.5, 8, 8),
    ("SO", 6, 8, "heat", 1.0, None, 1.0, 6, 7),
    ("SU", 2, 15, "heat", 0.8, None, 1.3, 14, 9),
    ("SU", 3, 12, "matern", 1.0, 0.5, 1.0, 10, 10),
    ("SU", 4, 9, "heat", 1.2, None, 1.0, 7, 9),
])
def test_eigensum_on_groups(ref, family, n, order, kind, ls, nu, var, nx, ny):
    torch.manual_seed(100 + n + order)
    og = orc.Group(family, n, order)
    om = orc.Measure(kind, og.dim, ls, nu, var)
    x, y = og.rand(nx), og.rand(ny)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        space = getattr(ref["spaces"], family)(n, order=order)
        kern = ref["sk"].EigenbasisSumKernel(_ref_measure(ref, kind, og.dim, ls, nu, var), space)
    kern.eval()
    assert [tuple(e.index) for e in space.lb_eigenspaces] == [tuple(i[0]) for i in og.irreps]
    assert [e.dimension for e in space.lb_eigenspaces] == [i[1] for i in og.irreps]
    assert [float(e.lb_eigenvalue) for e in space.lb_eigenspaces] == [float(i[2]) for i in og.irreps]
    with torch.no_grad():
        want_raw = kern(x, y, normalize=False)
        want = kern(x, y)
    got_raw = orc.eigensum_unnormalised(og, om, x, y)
    assert rel_err(got_raw, want_raw) < TOL
    got = orc.eigensum(og, om, x, y, kern.normalizer.detach())
    assert rel_err(got, want) < TOL
    # the torus representatives themselves (same LAPACK calls, same ordering rules)
    emb_ref = space.pairwise_embed(x[:4], y[:5])
    emb = og.pairwise_embed(x[:4], y[:5])
    if family == "SO":
        assert scaled_err(torch.view_as_real(emb), torch.view_as_real(emb_ref)) < 1e-12


@pytest.mark.parametrize("cls,okind,n,m,order,avg", [
    ("Stiefel", "stiefel", 4, 2, 8, 6), ("Stiefel", "stiefel", 5, 3, 10, 4),
    ("Grassmannian", "grassmannian", 4, 2, 8, 6), ("OrientedGrassmannian", "oriented_grassmannian", 5, 2, 8, 5),
])
def test_homogeneous_spaces(ref, cls, okind, n, m, order, avg):
    torch.manual_seed(200 + n + m)
    oh = orc.Homogeneous(okind, n, m, order=order, average_order=avg)
    om = orc.Measure("matern", oh.dim, 0.9, 2.5, 1.4)
    x, y, phases = oh.rand(7), oh.rand(6), oh.rand(5)
    weights = torch.randn(5 * order, dtype=torch.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        space = getattr(ref["spaces"], cls)(n, m, order=order, average_order=avg)
        space.h_samples = oh.h_samples.clone()
        meas = _ref_measure(ref, "matern", oh.dim, 0.9, 2.5, 1.4)
        es = ref["sk"].EigenbasisSumKernel(meas, space)
        rp = ref["sk"].RandomPhaseKernel(meas, space, phase_order=5)
    es.eval()
    rp.eval()
    rp.phases = phases.clone()
    assert space.dim == oh.dim
    with torch.no_grad():
        assert scaled_err(oh.lift(x), space.M_to_G(x)) < 1e-13
        assert cov_err(orc.eigensum_unnormalised(oh, om, x, y), es(x, y, normalize=False)) < TOL
        emb_ref = rp.make_embedding(x)
        assert scaled_err(orc.random_phase_embedding(oh, om, phases, x, torch.sqrt), emb_ref) < TOL
        rp.compute_normalizer()
        want = rp(x, y).evaluate()
        assert cov_err(orc.random_phase_kernel(oh, om, phases, x, y, rp.normalizer.detach()), want) < TOL
        sampler = ref["pa"].RandomPhaseApproximation(es, phase_order=5)
        sampler.phases, sampler.weights = phases.clone(), weights.clone()
        assert scaled_err(orc.random_phase_sample(oh, om, phases, weights, x, es.normalizer.detach()), sampler(x)) < TOL


@pytest.mark.parametrize("name,n,kind,nu", [("hyp", 2, "heat", None), ("hyp", 3, "matern", 2.5), ("hyp", 4, "matern", 1.5),
                                            ("spd", 2, "matern", 2.5), ("spd", 3, "heat", None)])
def test_noncompact_fourier_features(ref, name, n, kind, nu):
    torch.manual_seed(300 + n)
    m = 24
    osp = orc.Hyperbolic(n, order=m) if name == "hyp" else orc.SPD(n, order=m)
    om = orc.Measure(kind, osp.dim, 0.8, nu, 1.6)
    x, y = osp.rand(9), osp.rand(8)
    lmd, shift = osp.generate(om)
    w_re, w_im = torch.randn(m, dtype=torch.float64), torch.randn(m, dtype=torch.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls = ref["spaces"].HyperbolicSpace if name == "hyp" else ref["spaces"].SymmetricPositiveDefiniteMatrices
        space = cls(n, order=m)
        meas = _ref_measure(ref, kind, osp.dim, 0.8, nu, 1.6)
        kern = ref["sk"].RandomFourierFeatureKernel(meas, space)
    kern.eval()
    # inject the oracle's spectral samples and shifts into the reference's feature object
    funcs = space.lb_eigenspaces
    funcs.exp.lmd, funcs.exp.shift = lmd.clone(), shift.clone()
    funcs.coeff = funcs.c_function(lmd) if name == "hyp" else funcs.c_function_tanh(lmd)     # hyperbolic.py:134-150 / spd.py:116-123
    assert scaled_err(osp.c_function(lmd), funcs.coeff.reshape(-1)) < 1e-13
    with torch.no_grad():
        f_ref = funcs(space.to_group(x))
        f = osp.features(osp.to_group(x), lmd, shift)
        assert scaled_err(torch.view_as_real(f), torch.view_as_real(f_ref)) < TOL
        kern.compute_normalizer()
        norm = orc.rff_normalizer(osp, lmd, shift)
        assert abs(float(norm) / float(kern.normalizer) - 1) < 1e-12
        assert cov_err(orc.rff_kernel(osp, om, lmd, shift, x, y, norm), kern(x, y).evaluate()) < TOL
        sampler = ref["pa"].RandomFourierApproximation(kern)
        sampler.weights_real, sampler.weights_imag = w_re.clone(), w_im.clone()
        assert scaled_err(orc.rff_sample(osp, om, lmd, shift, w_re, w_im, x, float(norm)), sampler(x)) < TOL


def test_rand_consumes_the_generator_like_the_reference(ref):
    """`space.rand` of the oracle draws the same numbers in the same order as the reference (a1-a4 of SURVEY.md 8a)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cases = [(ref["spaces"].SO(3, order=2), orc.Group("SO", 3, 2)), (ref["spaces"].SO(5, order=2), orc.Group("SO", 5, 2)),
                 (ref["spaces"].SU(2, order=2), orc.Group("SU", 2, 2)), (ref["spaces"].SU(4, order=2), orc.Group("SU", 4, 2)),
                 (ref["spaces"].HyperbolicSpace(3, order=4), orc.Hyperbolic(3, order=4)),
                 (ref["spaces"].SymmetricPositiveDefiniteMatrices(3, order=4), orc.SPD(3, order=4))]
    for rs, os_ in cases:
        torch.manual_seed(77)
        a = rs.rand(6)
        torch.manual_seed(77)
        b = os_.rand(6)
        a_r = torch.view_as_real(a) if torch.is_complex(a) else a
        b_r = torch.view_as_real(b) if torch.is_complex(b) else b
        assert torch.equal(a_r, b_r), type(rs).__name__
