"""GPU parity for RandomPhaseKernel / RandomPhaseApproximation on group manifolds and the fp64 DMMA Gram."""
import pytest
import torch

from conftest import load_golden, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _measure(spec):
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    if spec["kind"] == "matern":
        return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], spec.get("variance", 1.0))
    return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], spec.get("variance", 1.0))


@pytest.mark.parametrize("n,m,k", [(1, 1, 1), (70, 33, 19), (128, 64, 32), (200, 131, 257), (513, 300, 1000)])
def test_gram_matches_torch(n, m, k):
    from liestationarykernels_b200 import ops
    torch.manual_seed(0)
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    b = torch.randn(m, k, dtype=torch.float64, device="cuda")
    pa, pb = ops.PanelMatrix.from_dense(a), ops.PanelMatrix.from_dense(b)
    assert torch.equal(pa.to_dense(), a)
    got = ops.gram(pa, pb, 0.5)
    ref = 0.5 * a @ b.T
    assert (got - ref).abs().max().item() < 1e-12 * k ** 0.5 * 10


@pytest.mark.parametrize("name", ["rpk_so3", "rpk_so5", "rpk_su3"])
def test_random_phase_on_groups(name):
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomPhaseKernel
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    g = load_golden(name)
    space = (SO if g["group"] == "SO" else SU)(g["n"], order=g["order"])
    meas = _measure(g["measure"])
    P = g["phases"].shape[0]
    kern = RandomPhaseKernel(meas, space, phase_order=P)
    kern.eval()
    kern.phases = g["phases"].cuda()
    kern.normalizer = g["normalizer"].cuda()
    x, y = g["x"].cuda(), g["y"].cuda()
    with torch.no_grad():
        emb = kern.make_embedding(x).cpu()
        assert emb.shape == g["embedding"].shape
        assert scaled_err(emb, g["embedding"]) < TOL
        gram = kern(x, y).evaluate().cpu()
        assert scaled_err(gram, g["gram"]) < TOL
        esum = EigenbasisSumKernel(meas, space)
        esum.eval()
        esum.normalizer = g["es_normalizer"].cuda()
        sampler = RandomPhaseApproximation(esum, phase_order=P)
        sampler.phases = g["s_phases"].cuda()
        sampler.weights = g["s_weights"].cuda()
        sample = sampler(x).cpu()
        assert scaled_err(sample, g["sample"]) < TOL
