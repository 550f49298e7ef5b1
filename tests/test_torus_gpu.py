"""GPU parity of the flat torus: EigenbasisSumKernel through lsk_torus_features + the Gram kernel
against golden outputs of the real reference, the oracle port on ragged sizes, and hyper-parameter gradients."""
import pytest
import torch

from conftest import cov_err, load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-10     # of the kernel's scale: truncated character sums on the torus cross zero


def _measure(spec):
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    if spec["kind"] == "matern":
        return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], spec.get("variance", 1.0))
    return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], spec.get("variance", 1.0))


@pytest.mark.parametrize("name", ["eigsum_torus2_matern_L20", "eigsum_torus3_heat_L31", "eigsum_torus1_heat_L15"])
def test_golden(name):
    from liestationarykernels_b200.spaces import Torus
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    g = load_golden(name)
    space = Torus(g["n"], order=g["order"])
    assert [(tuple(e.index), e.dimension, e.lb_eigenvalue) for e in space.lb_eigenspaces] == [(tuple(s), d, c) for s, d, c in g["irreps"]]
    kern = EigenbasisSumKernel(_measure(g["measure"]), space)
    kern.eval()
    x, y = g["x"].cuda(), g["y"].cuda()
    with torch.no_grad():
        raw = kern(x, y, normalize=False).cpu()
        assert cov_err(raw, g["k_raw"]) < TOL
        big = torch.abs(g["k_raw"]) > 1e-3 * torch.abs(g["k_raw"]).max()           # per-entry bound away from the zero crossings
        assert (torch.abs(raw - g["k_raw"])[big] / torch.abs(g["k_raw"])[big]).max().item() < 1e-9
        assert abs(float(kern.normalizer) / float(g["normalizer"]) - 1) < 1e-12
        kern.normalizer = g["normalizer"].cuda()
        assert cov_err(kern(x, y).cpu(), g["k_norm"]) < TOL
        assert cov_err(kern(x[:8]).cpu(), g["k_xx"]) < TOL
        assert torch.allclose(space.pairwise_dist(x[:16], y[:12]).cpu(), g["pairwise_dist"], rtol=1e-14, atol=0)
        assert torch.allclose(space.dist(x[:16], y[:16]).cpu(), g["dist"], rtol=1e-14, atol=0)
        # the character modules the reference exposes
        e = space.lb_eigenspaces[3]
        d = space.pairwise_embed(x[:5], y[:4])
        k = torch.tensor(e.index, dtype=torch.float64, device="cuda")
        assert torch.allclose(e.phase_function(d).real, torch.cos(2 * 3.1415927410125732 * (d * k).sum(1)), atol=1e-14)


def test_against_oracle_ragged_and_gradients():
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import Torus
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(11)
    ot = orc.Torus(2, 25)
    x, y = ot.rand(131), ot.rand(67)
    om = orc.Measure("matern", 2, 0.25, 1.5, 1.4)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    emb = ot.pairwise_embed(x, y)
    cov = torch.zeros(131, 67, dtype=torch.float64)
    norm = torch.zeros((), dtype=torch.float64)
    for irrep in ot.irreps:
        cov = cov + om(irrep[2]) * ot.phase_function(irrep, emb).view(131, 67).real
        norm = norm + om(irrep[2])[0]
    k_ref = torch.abs(om.variance[0]) * cov / norm
    g_out = torch.randn(131, 67, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    meas = MaternSpectralMeasure(2, 0.25, 1.5, 1.4)
    kern = EigenbasisSumKernel(meas, Torus(2, order=25))
    kern.train()
    k = kern(x.cuda(), y.cuda())
    assert cov_err(k.detach().cpu(), k_ref.detach()) < TOL
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert abs(ours.item() - ref.item()) < 1e-9 * max(1.0, abs(ref.item())), (ours.item(), ref.item())
    kern.eval()
    with torch.no_grad():
        assert kern(x[:0].cuda(), y.cuda()).shape == (0, 67)
