"""GPU parity for RandomPhaseKernel / RandomPhaseApproximation on group manifolds and the fp64 DMMA Gram."""
import pytest
import torch

from conftest import cov_err, load_golden, scaled_err

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
        assert cov_err(gram, g["gram"]) < TOL
        esum = EigenbasisSumKernel(meas, space)
        esum.eval()
        esum.normalizer = g["es_normalizer"].cuda()
        sampler = RandomPhaseApproximation(esum, phase_order=P)
        sampler.phases = g["s_phases"].cuda()
        sampler.weights = g["s_weights"].cuda()
        sample = sampler(x).cpu()
        assert scaled_err(sample, g["sample"]) < TOL


@pytest.mark.parametrize("case,P", [("so3", 32), ("so3", 30), ("stiefel53", 16), ("sphere3", 24)])
def test_random_phase_kernel_hyperparameter_gradients(case, P):
    """RandomPhaseKernel in training mode (the kernel `examples/gpr_experiments.py:36-39,66-67` trains on homogeneous spaces):
    value and lengthscale / variance gradients against torch autograd through the oracle port, normaliser recomputed at the
    identity from the same parameters (spectral_kernel.py:89-91); P = 30 exercises the unaligned column-block path."""
    import lsk_oracle as orc
    from liestationarykernels_b200.spaces import SO, Sphere, Stiefel
    from liestationarykernels_b200.spectral_kernel import RandomPhaseKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    torch.manual_seed(17)
    if case == "so3":
        osp, space = orc.Group("SO", 3, 10), SO(3, order=10)
    elif case == "stiefel53":
        osp = orc.Homogeneous("stiefel", 5, 3, order=10, average_order=12)
        space = Stiefel(5, 3, order=10, average_order=12)
        space.h_samples = osp.h_samples.cuda()
    else:
        osp, space = orc.Sphere(3, 10), Sphere(3, order=10)
    nx, ny = 37, 21
    x, y, phases = osp.rand(nx), osp.rand(ny), osp.rand(P)
    om = orc.Measure("matern", osp.dim, 0.7, 1.5, 1.9)
    om.lengthscale.requires_grad_(True)
    om.variance.requires_grad_(True)
    norm = orc.random_phase_kernel(osp, om, phases, osp.id, osp.id)[0, 0]
    k_ref = orc.random_phase_kernel(osp, om, phases, x, y, norm)
    g_out = torch.randn(nx, ny, dtype=torch.float64)
    (k_ref * g_out).sum().backward()
    meas = MaternSpectralMeasure(osp.dim, 0.7, 1.5, 1.9)
    kern = RandomPhaseKernel(meas, space, phase_order=P)
    kern.phases = phases.cuda()
    kern.train()
    k = kern(x.cuda(), y.cuda()).evaluate()
    assert cov_err(k.detach().cpu(), k_ref.detach()) < TOL
    assert abs(float(kern.normalizer) - norm.item()) < 1e-10 * abs(norm.item())
    (k * g_out.cuda()).sum().backward()
    for ours, ref in ((meas.lengthscale.grad, om.lengthscale.grad), (meas.variance.grad, om.variance.grad)):
        assert ours is not None
        assert abs(ours.item() - ref.item()) < 1e-8 * max(1.0, abs(ref.item())), (ours.item(), ref.item())
    # eval mode: stored normaliser, still differentiable; no_grad: the plain lazy path gives the same numbers
    kern.eval()
    k2 = kern(x.cuda(), y.cuda()).evaluate()
    with torch.no_grad():
        k3 = kern(x.cuda(), y.cuda()).evaluate()
    assert scaled_err(k2.detach().cpu(), k3.cpu()) < 1e-13
    assert k2.requires_grad and not k3.requires_grad


@pytest.mark.parametrize("n,k", [(1, 1), (70, 19), (128, 64), (513, 1000), (2049, 257)])
def test_panel_matvec_and_column_block_gram(n, k):
    """lsk_panel_gemv and the column-block Gram (operand pointers advanced by whole k-groups) against dense torch"""
    from liestationarykernels_b200 import ops
    torch.manual_seed(n + k)
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    w = torch.randn(k, dtype=torch.float64, device="cuda")
    pa = ops.PanelMatrix.from_dense(a)
    got = ops.panel_matvec(pa, w, 0.75)
    ref = 0.75 * (a @ w)
    assert (got - ref).abs().max().item() <= 1e-13 * max(1.0, ref.abs().max().item()) * k ** 0.5
    assert torch.equal(got, ops.panel_matvec(pa, w, 0.75))                      # deterministic
    if k >= 8:
        m = max(n // 2, 1)
        b = torch.randn(m, k, dtype=torch.float64, device="cuda")
        pb = ops.PanelMatrix.from_dense(b)
        g0, g1 = 1, min(pa.kg, 1 + max((k // 4) // 2, 1))
        blk = ops.gram_columns(pa, pb, g0, g1, 2.0)
        c0, c1 = 4 * g0, min(4 * g1, k)
        ref_blk = 2.0 * a[:, c0:c1] @ b[:, c0:c1].T
        assert (blk - ref_blk).abs().max().item() <= 1e-12 * max(1.0, ref_blk.abs().max().item())


@pytest.mark.parametrize("n,k,ld_pad", [(64, 128, 0), (200, 128, 8), (333, 256, 1), (1000, 64, 0), (1111, 128, 6)])
def test_rank_k_update_epilogue(n, k, ld_pad):
    """SYRK kind of lsk_pairwise_poly: out += scale * A B^T in place, rectangular and on the triangular tile schedule (tiles
    that reach the diagonal or lie right of it), ragged sizes, padded / odd leading dimension (scalar-store path)."""
    import ctypes
    from liestationarykernels_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(n)
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    b = torch.randn(n - n // 3, k, dtype=torch.float64, device="cuda")
    pa, pb = ops.PanelMatrix.from_dense(a), ops.PanelMatrix.from_dense(b)
    ld = n + ld_pad
    for tri in (0, 1):
        rows, cols, right = (n, n, pa) if tri else (n, b.shape[0], pb)
        full = torch.randn(rows, ld, dtype=torch.float64, device="cuda")
        before = full.clone()
        ep = _lib.LskEpilogue()
        ep.kind, ep.nforms, ep.nvars, ep.nterms = _lib.EPI_SYRK, 1, 1, 1
        ep.group_begin[0], ep.group_end[0] = 0, pa.kg
        ep.scale, ep.out_layout, ep.reserved = -1.5, _lib.OUT_ROW_MAJOR, 2 * tri
        rc = lib.lsk_pairwise_poly(_lib.ptr(pa.buf), _lib.ptr(right.buf), rows, cols, pa.kg, ctypes.byref(ep), _lib.ptr(full), ld,
                                   _lib.stream_ptr())
        assert rc == 0
        want = before[:, :cols] - 1.5 * a @ (a if tri else b).T
        diff = (full[:, :cols] - want).abs()
        if tri:
            # entries in tiles that reach the diagonal are updated (128-row x 64-column tiles); below them the matrix is untouched
            ii = torch.arange(rows, device="cuda")[:, None]
            jj = torch.arange(cols, device="cuda")[None, :]
            touched = (jj // 64) >= 2 * (ii // 128)
            assert diff[touched].max().item() <= 1e-12 * k ** 0.5
            assert torch.equal(full[:, :cols][~touched], before[:, :cols][~touched])
        else:
            assert diff.max().item() <= 1e-12 * k ** 0.5
        assert torch.equal(full[:, cols:], before[:, cols:])                    # padding columns untouched


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_symmetric_gram_reassembles_the_upper_triangle(world, monkeypatch):
    """`distributed.sharded_symmetric_gram` for every rank of an emulated world on one GPU: row blocks of one panel-major feature map
    addressed in place (`PanelMatrix.row_block`), diagonal blocks on the triangular tile schedule, rectangles in general mode.
    Each entry of the upper triangle must be bit-identical to the single-GPU symmetric-mode Gram."""
    from liestationarykernels_b200 import distributed as D, ops
    torch.manual_seed(world)
    n, k = 3000, 260
    a = torch.randn(n, k, dtype=torch.float64, device="cuda")
    phi = ops.PanelMatrix.from_dense(a)
    want = ops.gram(phi, phi, 0.5)
    assert torch.equal(phi.row_block(1280, 2000).to_dense(), a[1280:2000])
    seen = torch.zeros(n, n, dtype=torch.bool, device="cuda")
    for rank in range(world):
        monkeypatch.setattr(D, "world", lambda ws=world, rk=rank: (ws, rk))
        for lo, hi, block in D.sharded_symmetric_gram(phi, 0.5):
            assert block.shape == (hi - lo, n - lo)
            assert torch.equal(block, want[lo:hi, lo:])
            assert not seen[lo:hi, lo:].any()
            seen[lo:hi, lo:] = True
    assert bool((seen | seen.T).all())
