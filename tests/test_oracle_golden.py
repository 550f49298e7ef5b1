"""Pins the oracle port (oracle/lsk_oracle.py) against golden outputs of the REAL reference
(tests/golden/*.pt, produced by oracle/make_golden.py).  CPU only."""
import math

import pytest
import torch

from conftest import cov_err, load_golden, rel_err, scaled_err
import lsk_oracle as orc

TOL = 1e-11   # port vs reference: same algorithm, same LAPACK; differences are summation-order rounding


def _measure(spec):
    return orc.Measure(spec["kind"], spec["dim"], spec["lengthscale"], spec.get("nu"), spec.get("variance", 1.0))


EIG = ["eigsum_so3_heat_L20", "eigsum_so3_matern_L100", "eigsum_so5_heat_L20", "eigsum_so7_heat_L20",
       "eigsum_su2_heat_L20", "eigsum_su3_matern_L20", "eigsum_su4_heat_L20", "eigsum_su5_heat_L10",
       "eigsum_so4_heat_L20", "eigsum_so6_heat_L20", "eigsum_su6_heat_L10", "eigsum_su6_matern_L20",
       "eigsum_so8_heat_L20", "eigsum_so8_matern_L10"]


@pytest.mark.parametrize("name", EIG)
def test_eigensum_matches_reference(name):
    g = load_golden(name)
    space = orc.Group(g["group"], g["n"], g["order"])
    assert [(tuple(s), d, c) for s, d, c in space.irreps] == [(tuple(s), d, c) for s, d, c in g["irreps"]]
    meas = _measure(g["measure"])
    nx, ny = (16, 12) if g["order"] > 50 else (len(g["x"]), len(g["y"]))
    raw = orc.eigensum_unnormalised(space, meas, g["x"][:nx], g["y"][:ny])
    assert cov_err(raw, g["k_raw"][:nx, :ny]) < TOL
    full = orc.eigensum(space, meas, g["x"][:nx], g["y"][:ny], g["normalizer"])
    assert cov_err(full, g["k_norm"][:nx, :ny]) < TOL


@pytest.mark.parametrize("name", ["eigsum_so5_matern25_L100", "eigsum_so5_matern05_L100"])
def test_eigensum_so5_L100_block(name):
    g = load_golden(name)
    space = orc.Group("SO", 5, 100)
    meas = _measure(g["measure"])
    raw = orc.eigensum_unnormalised(space, meas, g["x"][:6], g["y"][:6])
    assert rel_err(raw, g["k_raw"][:6, :6]) < TOL


def test_characters_and_torus():
    rec = load_golden("chars_small")
    for key, r in rec.items():
        space = orc.Group(key[:2], int(key[2:]), 10)
        gam = space.torus_representative(r["x"])
        for (sig, _, _), ref in zip(space.irreps, r["chi"]):
            got = space.chi(sig, gam)
            assert torch.abs(got - ref).max().item() < 1e-10 * max(1.0, torch.abs(ref).max().item())


def test_rand_replays_reference_rng():
    rec = load_golden("rand_draws")
    for key, r in rec.items():
        torch.manual_seed(99)
        if key == "H3":
            got = orc.Hyperbolic(3, 4).rand(32)
            ref = r["x"]
        elif key == "SPD3":
            got = orc.SPD(3, 4).rand(32)
            ref = r["x"]
        else:
            got = orc.Group(key[:2], int(key[2:]), 0).rand(len(r["q"]))
            ref = r["q"]
        assert torch.equal(got, ref), key


HOM = {"homog_stiefel53": "stiefel", "homog_stiefel52_heat": "stiefel", "homog_grassmannian52": "grassmannian",
       "homog_orgrassmannian52": "oriented_grassmannian", "homog_stiefel32": "stiefel", "homog_stiefel42": "stiefel",
       "homog_grassmannian42": "grassmannian", "homog_orgrassmannian42": "oriented_grassmannian",
       "homog_grassmannian31": "grassmannian"}


@pytest.mark.parametrize("name", list(HOM))
def test_homogeneous_matches_reference(name):
    g = load_golden(name)
    n, m = g["nm"]
    nx = 8 if name == "homog_stiefel53" else len(g["x"])
    space = orc.Homogeneous(HOM[name], n, m, order=g["kw"]["order"], average_order=g["kw"]["average_order"])
    space.h_samples = g["h_samples"]
    meas = _measure(g["measure"])
    assert torch.abs(space.lift(g["x"]) - g["lift_x"]).max().item() < 1e-14
    emb = orc.random_phase_embedding(space, meas, g["phases"], g["x"][:nx], torch.sqrt)
    assert scaled_err(emb, g["embedding"][:nx]) < TOL
    gram = orc.random_phase_kernel(space, meas, g["phases"], g["x"][:nx], g["y"][:nx], g["normalizer"])
    assert cov_err(gram, g["gram"][:nx, :nx]) < TOL
    sample = orc.random_phase_sample(space, meas, g["s_phases"], g["s_weights"], g["x"][:nx], g["es_normalizer"])
    assert scaled_err(sample, g["sample"][:nx]) < TOL
    if nx <= 12:
        kes = orc.eigensum(space, meas, g["x"][:6], g["y"][:6], g["es_normalizer"])
        assert cov_err(kes, g["k_es"][:6, :6]) < TOL


@pytest.mark.parametrize("name", ["rpk_so3", "rpk_so5", "rpk_su3"])
def test_random_phase_on_groups(name):
    g = load_golden(name)
    space = orc.Group(g["group"], g["n"], g["order"])
    meas = _measure(g["measure"])
    emb = orc.random_phase_embedding(space, meas, g["phases"], g["x"], torch.sqrt)
    assert scaled_err(emb, g["embedding"]) < TOL
    gram = orc.random_phase_kernel(space, meas, g["phases"], g["x"], g["y"], g["normalizer"])
    assert cov_err(gram, g["gram"]) < TOL
    sample = orc.random_phase_sample(space, meas, g["s_phases"], g["s_weights"], g["x"], g["es_normalizer"])
    assert scaled_err(sample, g["sample"]) < TOL


RFF = ["rff_h3_matern", "rff_h3_heat", "rff_h2_heat", "rff_h5_matern", "rff_spd3_matern", "rff_spd3_heat",
       "rff_spd2_heat", "rff_spd4_heat"]


@pytest.mark.parametrize("name", RFF)
def test_noncompact_matches_reference(name):
    g = load_golden(name)
    space = orc.Hyperbolic(g["n"], g["order"]) if g["cls"] == "HyperbolicSpace" else orc.SPD(g["n"], g["order"])
    meas = _measure(g["measure"])
    assert torch.abs(space.c_function(g["lmd"]) - g["coeff"]).max().item() < 1e-13 * max(1.0, g["coeff"].abs().max().item())
    feats = space.features(space.to_group(g["x"]), g["lmd"], g["shift"])
    assert scaled_err(feats, g["features"]) < TOL
    norm = orc.rff_normalizer(space, g["lmd"], g["shift"])
    assert abs(norm.item() - g["normalizer"].item()) < 1e-12 * abs(g["normalizer"].item())
    gram = orc.rff_kernel(space, meas, g["lmd"], g["shift"], g["x"], g["y"], g["normalizer"])
    assert cov_err(gram, g["gram"]) < TOL
    sample = orc.rff_sample(space, meas, g["lmd"], g["shift"], g["w_re"], g["w_im"], g["x"], g["normalizer"])
    assert scaled_err(sample, g["sample"]) < TOL


@pytest.mark.parametrize("name", ["eigsum_torus2_matern_L20", "eigsum_torus3_heat_L31", "eigsum_torus1_heat_L15"])
def test_torus_eigensum_matches_reference(name):
    g = load_golden(name)
    space = orc.Torus(g["n"], g["order"])
    assert [(tuple(s), d, c) for s, d, c in space.irreps] == [(tuple(s), d, c) for s, d, c in g["irreps"]]
    meas = _measure(g["measure"])
    raw = orc.eigensum_unnormalised(space, meas, g["x"], g["y"])
    assert cov_err(raw, g["k_raw"]) < TOL
    assert cov_err(orc.eigensum(space, meas, g["x"], g["y"], g["normalizer"]), g["k_norm"]) < TOL
    assert torch.equal(space.pairwise_dist(g["x"][:16], g["y"][:12]), g["pairwise_dist"])
    assert torch.equal(space.dist(g["x"][:16], g["y"][:16]), g["dist"])


SPHERE_CASES = ["eigsum_sphere2_heat_L15", "eigsum_sphere3_matern_L10", "eigsum_sphere4_matern_L30", "eigsum_proj2_heat_L10",
                "eigsum_proj3_matern_L8"]


@pytest.mark.parametrize("name", SPHERE_CASES)
def test_sphere_eigensum_matches_reference(name):
    """oracle port of spaces/sphere.py against golden outputs of the real reference (kernel, prior sample, embedding)"""
    g = load_golden(name)
    space = orc.Sphere(g["n"], g["order"], projective=g["group"] == "ProjectiveSpace")
    assert space.irreps == [tuple(i) for i in g["irreps"]]
    meas = _measure(g["measure"])
    assert cov_err(orc.eigensum_unnormalised(space, meas, g["x"], g["y"]), g["k_raw"]) < TOL
    assert cov_err(orc.eigensum(space, meas, g["x"], g["y"], g["normalizer"]), g["k_norm"]) < TOL
    assert torch.allclose(space.pairwise_embed(g["x"][:6], g["y"][:5]), g["pairwise_embed"], rtol=0, atol=1e-15)
    # arccos is ill-conditioned at coinciding points (the first three): compare away from them
    assert torch.allclose(space.pairwise_dist(g["x"][:16], g["y"][:12])[3:], g["pairwise_dist"][3:], rtol=0, atol=1e-13)
    sample = orc.random_phase_sample(space, meas, g["phases"], g["weights"], g["x"], g["normalizer"])
    # degrees 20-24 on S^4 go through the reference's monomial Gegenbauer form (sphere.py:170-203), whose cancellation
    # amplifies the rounding of the summation order (vmap(dot) there, a batched sum here) to a few 1e-10
    assert scaled_err(sample, g["sample"]) < (2e-9 if name == "eigsum_sphere4_matern_L30" else TOL)


@pytest.mark.parametrize("n,order,projective", [(2, 15, False), (3, 24, False), (4, 40, False), (2, 12, True), (5, 10, True)])
def test_zonal_chebyshev_plan_matches_gegenbauer(n, order, projective):
    """host tables of spaces/sphere.py (no GPU needed): every row of the Chebyshev re-expansion, evaluated by Clenshaw in
    float64, against a 50-digit evaluation of d_l C_l(t) / C_l(1); the oracle's restatement of the reference's monomial
    evaluation is held to the same truth to document how far the REFERENCE is from it.  Degrees >= 25 follow the
    reference's `NewGegenbauerPolynomials` (recurrence seeded with C_0 = 0, sphere.py:209-217), quirk G10."""
    import mpmath
    import numpy as np
    from liestationarykernels_b200.spaces.sphere import _ZonalPlan, num_harmonics
    degrees = [(2 if projective else 1) * i for i in range(order)]
    dims = [num_harmonics(n + 1, d) for d in degrees]
    plan = _ZonalPlan(n, degrees, dims)
    space = orc.Sphere(n, order, projective=projective)
    assert [i[1] for i in space.irreps] == dims
    ts = np.concatenate([np.linspace(-1, 1, 41), [0.999999, -0.3333]])
    mpmath.mp.dps = 50
    alpha = mpmath.mpf(n - 1) / 2
    worst_ours, worst_ref = 0.0, 0.0

    def geg(deg, t):                       # three-term recurrence at 50 digits; degrees >= 25: the reference's C_0 = 0 seed
        c0, c1 = mpmath.mpf(1 if deg < 25 else 0), 2 * alpha * t
        if deg == 0:
            return c0
        for m in range(2, deg + 1):
            c0, c1 = c1, (2 * t * (m + alpha - 1) * c1 - (m + 2 * alpha - 2) * c0) / m
        return c1

    for l, d in enumerate(degrees):
        row = plan.M[l]
        assert np.all(row[d + 1:] == 0)
        truth = np.array([float(dims[l] * geg(d, mpmath.mpf(float(t))) / geg(d, mpmath.mpf(1))) for t in ts])
        b1 = np.zeros_like(ts)
        b2 = np.zeros_like(ts)
        for k in range(d, 0, -1):
            b1, b2 = row[k] + 2 * ts * b1 - b2, b1
        ours = row[0] + ts * b1 - b2
        ref = space.phase_function(space.irreps[l], torch.tensor(ts)).numpy()
        worst_ours = max(worst_ours, np.abs(ours - truth).max() / dims[l])
        worst_ref = max(worst_ref, np.abs(ref - truth).max() / dims[l])
    assert worst_ours < 5e-14
    assert worst_ref < 1e-7          # the reference's monomial form loses digits with the degree; ours does not
