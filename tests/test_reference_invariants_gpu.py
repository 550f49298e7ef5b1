"""The reference's own invariant tests (tests/lie_groups.py), re-run against this implementation through the same
Python API: conjugation invariance (:53-63), character at identity = dimension (:65-70), Monte-Carlo orthonormality of
characters (:72-84, atol 5e-2), x x^{-1} = I (:108-110)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GROUPS = [("SO", 3), ("SO", 4), ("SO", 5), ("SO", 6), ("SO", 7), ("SU", 2), ("SU", 3), ("SU", 4), ("SU", 5)]


def _space(fam, n, order=10):
    from liestationarykernels_b200.spaces import SO, SU
    return (SO if fam == "SO" else SU)(n, order=order)


@pytest.mark.parametrize("fam,n", GROUPS)
def test_character_conjugation_invariance(fam, n):
    torch.manual_seed(0)
    space = _space(fam, n)
    num = 20
    gs, xs = space.rand(num), space.rand(num)
    conj = torch.matmul(torch.matmul(gs[:, None], xs[None, :]), space.inv(gs)[:, None])     # [num, num, n, n]
    for irrep in space.lb_eigenspaces:
        chi = irrep.phase_function
        ref = chi.evaluate(xs)
        got = chi.evaluate(conj)
        assert got.shape == (num, num)
        assert torch.allclose(got, ref[None, :].expand(num, num), atol=1e-8 * max(1, irrep.dimension))


@pytest.mark.parametrize("fam,n", GROUPS)
def test_character_at_identity_equals_dimension(fam, n):
    space = _space(fam, n)
    for irrep in space.lb_eigenspaces:
        val = irrep.phase_function.evaluate(space.id).item()
        assert abs(val.real - irrep.dimension) < 1e-8 * irrep.dimension and abs(val.imag) < 1e-8 * irrep.dimension


@pytest.mark.parametrize("fam,n", [("SO", 3), ("SO", 5), ("SU", 2)])
def test_characters_orthogonality(fam, n):
    torch.manual_seed(1)
    space = _space(fam, n, order=6)
    x = space.rand(40000)
    gam = space.torus_representative(x)
    vals = torch.stack([irrep.phase_function.chi(gam) for irrep in space.lb_eigenspaces])     # [L, S]
    gram = (vals @ vals.conj().T / x.shape[0])
    eye = torch.eye(len(space.lb_eigenspaces), dtype=gram.dtype, device="cuda")
    assert torch.allclose(gram, eye, atol=8e-2)      # Monte-Carlo; the reference uses 1e5 samples and 5e-2


@pytest.mark.parametrize("fam,n", GROUPS)
def test_sampler_inverse(fam, n):
    torch.manual_seed(2)
    space = _space(fam, n, order=0)
    x = space.rand(100)
    eye = torch.eye(n, dtype=x.dtype, device="cuda").expand(100, n, n)
    assert torch.allclose(torch.matmul(x, space.inv(x)), eye, atol=1e-12)


def test_kernel_value_is_a_class_function_of_the_difference():
    """the fused kernel must agree with the API-compat path: sum_l a_l d_l Re chi_l(torus_representative(x y^{-1}))"""
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
    torch.manual_seed(3)
    space = _space("SO", 5, order=12)
    meas = SqExpSpectralMeasure(10, 0.9)
    kern = EigenbasisSumKernel(meas, space)
    kern.eval()
    x, y = space.rand(16), space.rand(12)
    with torch.no_grad():
        fused = kern(x, y, normalize=False)
        gam = space.pairwise_embed(x, y)
        slow = torch.zeros(16 * 12, dtype=torch.float64, device="cuda")
        for irrep in space.lb_eigenspaces:
            slow += meas(irrep.lb_eigenvalue)[0] * irrep.phase_function(gam).real
    assert torch.allclose(fused, slow.view(16, 12), rtol=1e-9, atol=0)
