"""TEST INFRASTRUCTURE — generate the golden fixtures under tests/golden/ by running the REAL reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Every fixture stores the *inputs* (points, phases, weights, h_samples, spectral samples) and the reference's
*outputs*, because CPU and CUDA generators differ: parity tests inject the same tensors into both sides
(SURVEY.md §8c "same-tensor rule").  Files are small (float64 tensors, a few hundred KB in total).
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main(only=None):
    warnings.simplefilter("ignore")
    ref_shim.import_reference()
    from lie_stationary_kernels.spaces import (SO, SU, Stiefel, Grassmannian, OrientedGrassmannian, HyperbolicSpace,
                                               SymmetricPositiveDefiniteMatrices)
    from lie_stationary_kernels.spectral_kernel import (EigenbasisSumKernel, RandomPhaseKernel,
                                                        RandomFourierFeatureKernel, RandomSpectralKernel)
    from lie_stationary_kernels.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    from lie_stationary_kernels.prior_approximation import RandomPhaseApproximation, RandomFourierApproximation

    os.makedirs(OUT, exist_ok=True)

    def measure_of(spec):
        if spec["kind"] == "matern":
            return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], spec.get("variance", 1.0))
        return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], spec.get("variance", 1.0))

    def _detach(o):
        if isinstance(o, torch.Tensor):
            return o.detach().clone()
        if isinstance(o, dict):
            return {k: _detach(v) for k, v in o.items()}
        return o

    def save(name, obj):
        torch.save(_detach(obj), os.path.join(OUT, name + ".pt"))
        print("wrote %-40s %.1f KB" % (name, os.path.getsize(os.path.join(OUT, name + ".pt")) / 1024), flush=True)

    def want(name):
        return only is None or any(name.startswith(o) for o in only)

    # ------------------------------------------------------------------ EigenbasisSumKernel on groups
    eig_cases = [
        ("eigsum_so3_heat_L20", SO, 3, 20, dict(kind="heat", dim=3, lengthscale=1.0), 96, 80),
        ("eigsum_so3_matern_L100", SO, 3, 100, dict(kind="matern", dim=3, lengthscale=0.4, nu=0.5), 40, 40),
        ("eigsum_so5_matern25_L100", SO, 5, 100, dict(kind="matern", dim=10, lengthscale=1.0, nu=2.5), 24, 24),
        ("eigsum_so5_matern05_L100", SO, 5, 100, dict(kind="matern", dim=10, lengthscale=0.4, nu=0.5), 20, 20),
        ("eigsum_so5_heat_L20", SO, 5, 20, dict(kind="heat", dim=10, lengthscale=0.8, variance=2.5), 40, 36),
        ("eigsum_so7_heat_L20", SO, 7, 20, dict(kind="heat", dim=21, lengthscale=1.0), 20, 20),
        ("eigsum_su2_heat_L20", SU, 2, 20, dict(kind="heat", dim=3, lengthscale=1.0), 40, 40),
        ("eigsum_su3_matern_L20", SU, 3, 20, dict(kind="matern", dim=8, lengthscale=1.0, nu=1.5), 32, 32),
        ("eigsum_su4_heat_L20", SU, 4, 20, dict(kind="heat", dim=15, lengthscale=1.0), 40, 36),
        ("eigsum_su5_heat_L10", SU, 5, 10, dict(kind="heat", dim=24, lengthscale=1.0), 16, 16),
        ("eigsum_su6_heat_L10", SU, 6, 10, dict(kind="heat", dim=35, lengthscale=1.0), 14, 12),
        ("eigsum_su6_matern_L20", SU, 6, 20, dict(kind="matern", dim=35, lengthscale=1.5, nu=2.5), 12, 12),
        ("eigsum_so8_heat_L20", SO, 8, 20, dict(kind="heat", dim=28, lengthscale=1.0), 12, 12),
        ("eigsum_so8_matern_L10", SO, 8, 10, dict(kind="matern", dim=28, lengthscale=2.0, nu=1.5), 10, 12),
        ("eigsum_so4_heat_L20", SO, 4, 20, dict(kind="heat", dim=6, lengthscale=1.0), 24, 24),
        ("eigsum_so6_heat_L20", SO, 6, 20, dict(kind="heat", dim=15, lengthscale=1.0), 20, 20),
    ]
    for name, cls, n, order, mspec, nx, ny in eig_cases:
        if not want(name):
            continue
        t0 = time.time()
        torch.manual_seed(1234)
        space = cls(n, order=order)
        kern = EigenbasisSumKernel(measure_of(mspec), space)
        kern.eval()
        x, y = space.rand(nx), space.rand(ny)
        y[:3] = x[:3]                                 # exercise the diagonal (identity class)
        with torch.no_grad():
            k_norm = kern(x, y)
            k_raw = kern(x, y, normalize=False)
            k_xx = kern(x[:8])
        save(name, dict(group=cls.__name__, n=n, order=order, measure=mspec, x=x, y=y,
                        normalizer=kern.normalizer.detach().clone(), k_norm=k_norm, k_raw=k_raw, k_xx=k_xx,
                        irreps=[(e.index, e.dimension, e.lb_eigenvalue) for e in space.lb_eigenspaces]))
        print("   %.1fs" % (time.time() - t0), flush=True)

    # ------------------------------------------------------------------ EigenbasisSumKernel on the flat torus (spaces/torus.py)
    from lie_stationary_kernels.spaces.torus import Torus
    torus_cases = [
        ("eigsum_torus2_matern_L20", 2, 20, dict(kind="matern", dim=2, lengthscale=0.3, nu=2.5), 96, 70),
        ("eigsum_torus3_heat_L31", 3, 31, dict(kind="heat", dim=3, lengthscale=0.2, variance=1.7), 64, 64),
        ("eigsum_torus1_heat_L15", 1, 15, dict(kind="heat", dim=1, lengthscale=0.1), 50, 41),
    ]
    for name, n, order, mspec, nx, ny in torus_cases:
        if not want(name):
            continue
        torch.manual_seed(4321)
        space = Torus(n, order=order)
        kern = EigenbasisSumKernel(measure_of(mspec), space)
        kern.eval()
        x, y = space.rand(nx), space.rand(ny)
        y[:3] = x[:3]
        with torch.no_grad():
            k_norm = kern(x, y)
            k_raw = kern(x, y, normalize=False)
            k_xx = kern(x[:8])
        save(name, dict(group="Torus", n=n, order=order, measure=mspec, x=x, y=y,
                        normalizer=kern.normalizer.detach().clone(), k_norm=k_norm, k_raw=k_raw, k_xx=k_xx,
                        irreps=[(e.index, e.dimension, e.lb_eigenvalue) for e in space.lb_eigenspaces],
                        pairwise_dist=space.pairwise_dist(x[:16], y[:12]), dist=space.dist(x[:16], y[:16])))

    # ------------------------------------------------------------------ Sphere / ProjectiveSpace zonal kernels (spaces/sphere.py)
    from lie_stationary_kernels.spaces.sphere import Sphere, ProjectiveSpace
    sphere_cases = [
        ("eigsum_sphere2_heat_L15", Sphere, 2, 15, dict(kind="heat", dim=2, lengthscale=1.0, variance=5.0), 64, 50),
        ("eigsum_sphere3_matern_L10", Sphere, 3, 10, dict(kind="matern", dim=3, lengthscale=0.5, nu=1.5, variance=4.0), 48, 40),
        ("eigsum_sphere4_matern_L30", Sphere, 4, 30, dict(kind="matern", dim=4, lengthscale=0.6, nu=2.5), 40, 33),
        ("eigsum_proj2_heat_L10", ProjectiveSpace, 2, 10, dict(kind="heat", dim=2, lengthscale=0.7), 48, 40),
        ("eigsum_proj3_matern_L8", ProjectiveSpace, 3, 8, dict(kind="matern", dim=3, lengthscale=1.0, nu=2.5), 32, 32),
    ]
    for name, cls, n, order, mspec, nx, ny in sphere_cases:
        if not want(name):
            continue
        torch.manual_seed(2468)
        space = cls(n, order=order)
        kern = EigenbasisSumKernel(measure_of(mspec), space)
        kern.eval()
        x, y = space.rand(nx), space.rand(ny)
        y[:3] = x[:3]
        sampler = RandomPhaseApproximation(kern, phase_order=16)
        with torch.no_grad():
            k_norm = kern(x, y)
            k_raw = kern(x, y, normalize=False)
            k_xx = kern(x[:8])
            sample = sampler(x)
            emb = sampler.make_embedding(x[:10])
        save(name, dict(group=cls.__name__, n=n, order=order, measure=mspec, x=x, y=y,
                        normalizer=kern.normalizer.detach().clone(), k_norm=k_norm, k_raw=k_raw, k_xx=k_xx,
                        irreps=[(e.index, e.dimension, e.lb_eigenvalue) for e in space.lb_eigenspaces],
                        phases=sampler.phases, weights=sampler.weights, sample=sample, embedding=emb,
                        pairwise_dist=space.pairwise_dist(x[:16], y[:12]), pairwise_embed=space.pairwise_embed(x[:6], y[:5])))

    # ------------------------------------------------------------------ torus representatives / characters (lie_groups.py tests)
    if want("chars"):
        torch.manual_seed(7)
        rec = {}
        for cls, n in ((SO, 3), (SO, 5), (SO, 7), (SU, 2), (SU, 3), (SU, 4)):
            space = cls(n, order=10)
            x = space.rand(12)
            gam = space.torus_representative(x)
            vals = torch.stack([e.phase_function.chi(gam) for e in space.lb_eigenspaces])
            rec["%s%d" % (cls.__name__, n)] = dict(x=x, chi=vals, sigs=[e.index for e in space.lb_eigenspaces])
        save("chars_small", rec)

    # ------------------------------------------------------------------ space.rand (inputs = the Gaussian draws)
    if want("rand"):
        rec = {}
        for cls, n, num in ((SO, 3, 64), (SO, 5, 64), (SO, 8, 32), (SU, 2, 64), (SU, 4, 48), (SU, 6, 16)):
            space = cls(n, order=0)
            torch.manual_seed(99)
            q = space.rand(num)
            torch.manual_seed(99)
            if (cls is SO and n == 3) or (cls is SU and n == 2):
                gauss = torch.randn((num, 4), dtype=torch.float64)
            else:
                gauss = torch.randn((num, n, n), dtype=torch.float64 if cls is SO else torch.complex128)
            rec["%s%d" % (cls.__name__, n)] = dict(gauss=gauss, q=q)
        hyp = HyperbolicSpace(3, order=4)
        torch.manual_seed(99)
        pts = hyp.rand(32)
        torch.manual_seed(99)
        g = torch.randn(32, 3, dtype=torch.float64)
        u = torch.rand(32, dtype=torch.float64)
        rec["H3"] = dict(gauss=g, unif=u, x=pts)
        spd = SymmetricPositiveDefiniteMatrices(3, order=4)
        torch.manual_seed(99)
        pts = spd.rand(32)
        torch.manual_seed(99)
        g = torch.randn(32, 3, 3, dtype=torch.float64)
        rec["SPD3"] = dict(gauss=g, x=pts)
        save("rand_draws", rec)

    # ------------------------------------------------------------------ homogeneous spaces: lift, embedding, Gram, sampler
    hom_cases = [
        ("homog_stiefel53", Stiefel, (5, 3), dict(order=10, average_order=30), dict(kind="matern", dim=9, lengthscale=1.0, nu=2.5), 20, 12),
        ("homog_stiefel52_heat", Stiefel, (5, 2), dict(order=10, average_order=12), dict(kind="heat", dim=7, lengthscale=0.7), 12, 8),
        ("homog_grassmannian52", Grassmannian, (5, 2), dict(order=10, average_order=20), dict(kind="heat", dim=6, lengthscale=1.0), 12, 8),
        ("homog_orgrassmannian52", OrientedGrassmannian, (5, 2), dict(order=10, average_order=16), dict(kind="heat", dim=6, lengthscale=1.0), 12, 8),
        ("homog_stiefel32", Stiefel, (3, 2), dict(order=10, average_order=10), dict(kind="heat", dim=3, lengthscale=1.0), 12, 8),
        ("homog_stiefel42", Stiefel, (4, 2), dict(order=10, average_order=12), dict(kind="heat", dim=5, lengthscale=1.0), 12, 8),
        ("homog_grassmannian42", Grassmannian, (4, 2), dict(order=10, average_order=12), dict(kind="heat", dim=4, lengthscale=1.0), 12, 8),
        ("homog_orgrassmannian42", OrientedGrassmannian, (4, 2), dict(order=10, average_order=12), dict(kind="heat", dim=4, lengthscale=1.0), 12, 8),
        ("homog_grassmannian31", Grassmannian, (3, 1), dict(order=10, average_order=10), dict(kind="heat", dim=2, lengthscale=1.0), 12, 8),
    ]
    for name, cls, nm, kw, mspec, nx, nph in hom_cases:
        if not want(name):
            continue
        t0 = time.time()
        torch.manual_seed(4321)
        space = cls(*nm, **kw)
        meas = measure_of(mspec)
        kern = RandomPhaseKernel(meas, space, phase_order=nph)
        kern.eval()
        x = space.rand(nx).contiguous()
        y = space.rand(nx - 3).contiguous()
        with torch.no_grad():
            emb = kern.make_embedding(x)
            gram = kern(x, y).evaluate()
            gram_raw = kern(x, y, normalize=False).evaluate()
            esum = EigenbasisSumKernel(meas, space)
            esum.eval()
            k_es = esum(x, y)
            sampler = RandomPhaseApproximation(esum, phase_order=nph)
            sample = sampler(x)
            samp_emb = sampler.make_embedding(x)
        save(name, dict(cls=cls.__name__, nm=nm, kw=kw, measure=mspec, x=x, y=y, h_samples=space.h_samples,
                        phases=kern.phases.contiguous(), normalizer=kern.normalizer.detach().clone(), embedding=emb,
                        gram=gram, gram_raw=gram_raw,
                        lift_x=space.M_to_G(x), es_normalizer=esum.normalizer.detach().clone(), k_es=k_es,
                        s_phases=sampler.phases.contiguous(), s_weights=sampler.weights, sample=sample, s_embedding=samp_emb))
        print("   %.1fs" % (time.time() - t0), flush=True)

    # ------------------------------------------------------------------ RandomPhaseKernel / sampler on a group manifold
    for name, cls, n, order, mspec, nx, nph in [
        ("rpk_so3", SO, 3, 12, dict(kind="heat", dim=3, lengthscale=1.0), 24, 16),
        ("rpk_so5", SO, 5, 10, dict(kind="matern", dim=10, lengthscale=1.0, nu=2.5), 16, 12),
        ("rpk_su3", SU, 3, 10, dict(kind="heat", dim=8, lengthscale=1.0), 16, 12),
    ]:
        if not want(name):
            continue
        torch.manual_seed(2468)
        space = cls(n, order=order)
        meas = measure_of(mspec)
        kern = RandomPhaseKernel(meas, space, phase_order=nph)
        kern.eval()
        x, y = space.rand(nx), space.rand(nx - 5)
        with torch.no_grad():
            emb = kern.make_embedding(x)
            gram = kern(x, y).evaluate()
            esum = EigenbasisSumKernel(meas, space)
            esum.eval()
            sampler = RandomPhaseApproximation(esum, phase_order=nph)
            sample = sampler(x)
        save(name, dict(group=cls.__name__, n=n, order=order, measure=mspec, x=x, y=y, phases=kern.phases,
                        normalizer=kern.normalizer.detach().clone(), embedding=emb, gram=gram,
                        es_normalizer=esum.normalizer.detach().clone(), s_phases=sampler.phases,
                        s_weights=sampler.weights, sample=sample))

    # ------------------------------------------------------------------ non-compact: H^n and SPD(n)
    for name, cls, n, order, mspec, nx in [
        ("rff_h3_matern", HyperbolicSpace, 3, 96, dict(kind="matern", dim=3, lengthscale=1.0, nu=2.5), 40),
        ("rff_h3_heat", HyperbolicSpace, 3, 96, dict(kind="heat", dim=3, lengthscale=1.0), 40),
        ("rff_h2_heat", HyperbolicSpace, 2, 64, dict(kind="heat", dim=2, lengthscale=0.8), 24),
        ("rff_h5_matern", HyperbolicSpace, 5, 64, dict(kind="matern", dim=5, lengthscale=1.0, nu=1.5), 24),
        ("rff_spd3_matern", SymmetricPositiveDefiniteMatrices, 3, 96, dict(kind="matern", dim=6, lengthscale=1.0, nu=2.5), 40),
        ("rff_spd3_heat", SymmetricPositiveDefiniteMatrices, 3, 96, dict(kind="heat", dim=6, lengthscale=1.0), 40),
        ("rff_spd2_heat", SymmetricPositiveDefiniteMatrices, 2, 64, dict(kind="heat", dim=3, lengthscale=1.0), 24),
        ("rff_spd4_heat", SymmetricPositiveDefiniteMatrices, 4, 48, dict(kind="heat", dim=10, lengthscale=1.5), 16),
    ]:
        if not want(name):
            continue
        torch.manual_seed(1357)
        space = cls(n, order=order)
        meas = measure_of(mspec)
        kern = RandomFourierFeatureKernel(meas, space)
        kern.eval()
        x, y = space.rand(nx), space.rand(nx - 7)
        with torch.no_grad():
            feats = space.lb_eigenspaces(space.to_group(x))
            gram = kern(x, y).evaluate()
            sampler = RandomFourierApproximation(kern)
            sample = sampler(x)
            cov = sampler._cov(x, y)
        save(name, dict(cls=cls.__name__, n=n, order=order, measure=mspec, x=x, y=y,
                        lmd=space.lb_eigenspaces.exp.lmd, shift=space.lb_eigenspaces.exp.shift,
                        coeff=space.lb_eigenspaces.coeff, normalizer=kern.normalizer.detach().clone(),
                        features=feats, gram=gram, w_re=sampler.weights_real, w_im=sampler.weights_imag,
                        sample=sample, cov=cov))

    # ------------------------------------------------------------------ TranslatedCharactersBasis (space.py:284-314) on small irreps
    for name, cls, n, order in (("basis_so3", SO, 3, 3), ("basis_su2", SU, 2, 3), ("basis_su3", SU, 3, 3)):
        if not want(name):
            continue
        torch.manual_seed(777)
        space = cls(n, order=order)
        x, y = space.rand(9), space.rand(7)
        irreps = []
        for eig in space.lb_eigenspaces:
            basis = eig.basis                                   # draws dim^2 translations, factorises their Gram
            irreps.append(dict(index=eig.index, dimension=eig.dimension, translations=basis.translations, coeffs=basis.coeffs,
                               f_x=basis(x), f_y=basis(y)))
        save(name, dict(group=cls.__name__, n=n, order=order, x=x, y=y, irreps=irreps))

    # ------------------------------------------------------------------ GP regression step (examples/gpr_model.py:26-39, 59-66)
    # covariance and its hyper-parameter gradients from the REAL reference kernel (plain torch, differentiable); marginal likelihood
    # and posterior by the LAPACK restatement of gpytorch's dense path (oracle gp_*; gpytorch itself is absent: see lsk_oracle.py)
    import lsk_oracle as orc
    for name, cls, n, order, mspec in (("gp_so3_matern", SO, 3, 10, dict(kind="matern", dim=3, lengthscale=0.8, nu=1.5, variance=1.7)),
                                       ("gp_su3_heat", SU, 3, 8, dict(kind="heat", dim=8, lengthscale=1.1, variance=0.9))):
        if not want(name):
            continue
        torch.manual_seed(4242)
        space = cls(n, order=order)
        meas = measure_of(mspec)
        kern = EigenbasisSumKernel(meas, space)
        kern.train()
        n_train, n_test = 48, 12
        x, xs = space.rand(n_train), space.rand(n_test)
        targets = torch.randn(n_train, dtype=torch.float64)
        raw_noise = torch.tensor([0.3], dtype=torch.float64, requires_grad=True)
        const = torch.tensor([0.2], dtype=torch.float64, requires_grad=True)
        k_train = kern(x)                                              # training mode: normaliser recomputed (consumes RNG)
        loss = -orc.gp_log_marginal(k_train, orc.gp_noise(raw_noise)[0], const.expand(n_train), targets)
        loss.backward()
        kern.eval()
        with torch.no_grad():
            k_eval, k_star, k_ss = kern(x), kern(x, xs), kern(xs)
            mean, cov = orc.gp_posterior(k_eval, k_star, k_ss, orc.gp_noise(raw_noise)[0], const.detach(), targets)
        save(name, dict(group=cls.__name__, n=n, order=order, measure=mspec, x=x, xs=xs, targets=targets, raw_noise=raw_noise.detach(),
                        constant=const.detach(), k_train=k_train, loss=loss, grad_lengthscale=meas.lengthscale.grad,
                        grad_variance=meas.variance.grad, grad_raw_noise=raw_noise.grad, grad_constant=const.grad,
                        eval_normalizer=kern.normalizer, posterior_mean=mean, posterior_cov=cov))

if __name__ == "__main__":
    main(sys.argv[1:] or None)
