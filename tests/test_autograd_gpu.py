"""Hyper-parameter gradients of EigenbasisSumKernel against the oracle port differentiated by
torch autograd on CPU (same formula as the reference: a_l(lengthscale) weights, normaliser = kernel at the identity)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group,n,order,kind", [("SO", 3, 10, "heat"), ("SO", 5, 20, "matern"), ("SU", 3, 10, "heat")])
def test_lengthscale_and_variance_gradients(group, n, order, kind):
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    torch.manual_seed(0)
    og = orc.Group(group, n, order)
    x, y = og.rand(24), og.rand(17)
    dim = og.dim
    # ---- oracle: differentiate the reference formula on CPU ----
    om = orc.Measure(kind, dim, 0.8, 1.5 if kind == "matern" else None, 1.7)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    emb = og.pairwise_embed(x, y)
    cov = torch.zeros(24, 17, dtype=torch.float64)
    norm = torch.zeros((), dtype=torch.float64)
    for irrep in og.irreps:
        cov = cov + om(irrep[2]) * og.phase_function(irrep, emb).view(24, 17).real
        norm = norm + om(irrep[2])[0] * irrep[1] ** 2
    k_ref = torch.abs(om.variance[0]) * cov / norm
    g_out = torch.randn(24, 17, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    # ---- ours ----
    space = (SO if group == "SO" else SU)(n, order=order)
    meas = MaternSpectralMeasure(dim, 0.8, 1.5, 1.7) if kind == "matern" else SqExpSpectralMeasure(dim, 0.8, 1.7)
    kern = EigenbasisSumKernel(meas, space)
    kern.train()
    k = kern(x.cuda(), y.cuda())
    assert (k.detach().cpu() - k_ref.detach()).abs().max().item() < 1e-10 * k_ref.abs().max().item()
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-9 * max(1.0, abs(ref.item())), (ours.item(), ref.item())


def test_adam_step_moves_hyperparameters():
    """the GP-regression use case of examples/gpr_model.py: an optimiser step through the kernel"""
    from liestationarykernels_b200.spaces import SO
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(0)
    space = SO(3, order=10)
    meas = MaternSpectralMeasure(3, 1.0, 2.5)
    kern = EigenbasisSumKernel(meas, space)
    x = space.rand(64)
    opt = torch.optim.Adam(kern.parameters(), lr=0.05)
    target = torch.eye(64, dtype=torch.float64, device="cuda")
    before = meas.lengthscale.item()
    for _ in range(3):
        opt.zero_grad()
        loss = ((kern(x, x) - target) ** 2).mean()
        loss.backward()
        opt.step()
    assert meas.lengthscale.item() != before


@pytest.mark.parametrize("kind,n,m", [("stiefel", 5, 3), ("grassmannian", 4, 2)])
def test_homogeneous_space_gradients(kind, n, m):
    """EigenbasisSumKernel on SO(n)/H in training mode: lengthscale / variance gradients against torch autograd through
    the oracle port (normaliser = kernel at coinciding points, recomputed from the same parameters)."""
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import Grassmannian, Stiefel
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(1)
    oh = orc.Homogeneous(kind, n, m, order=10, average_order=12)
    x, y = oh.rand(14), oh.rand(9)
    om = orc.Measure("matern", oh.dim, 0.9, 2.5, 1.3)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    emb = oh.pairwise_embed(x, y)
    emb_id = oh.pairwise_embed(oh.id, oh.id)
    cov = torch.zeros(14, 9, dtype=torch.float64)
    norm = torch.zeros((), dtype=torch.float64)
    for irrep in oh.irreps:
        cov = cov + om(irrep[2]) * oh.phase_function(irrep, emb).view(14, 9).real
        norm = norm + om(irrep[2])[0] * oh.phase_function(irrep, emb_id).real[0]
    k_ref = torch.abs(om.variance[0]) * cov / norm
    g_out = torch.randn(14, 9, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    space = (Stiefel if kind == "stiefel" else Grassmannian)(n, m, order=10, average_order=12)
    space.h_samples = oh.h_samples.cuda()                       # same-tensor rule: inject the oracle's subgroup samples
    meas = MaternSpectralMeasure(oh.dim, 0.9, 2.5, 1.3)
    kern = EigenbasisSumKernel(meas, space)
    kern.train()
    k = kern(x.cuda(), y.cuda())
    assert (k.detach().cpu() - k_ref.detach()).abs().max().item() < 1e-10 * k_ref.abs().max().item()
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-9 * max(1.0, abs(ref.item())), (ours.item(), ref.item())
