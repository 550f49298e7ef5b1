"""Pins the oracle port (oracle/lsk_oracle.py) against golden outputs of the REAL reference
(tests/golden/*.pt, produced by oracle/make_golden.py).  CPU only."""
import math

import pytest
import torch

from conftest import load_golden, rel_err, scaled_err
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
    assert scaled_err(raw, g["k_raw"][:nx, :ny]) < TOL
    full = orc.eigensum(space, meas, g["x"][:nx], g["y"][:ny], g["normalizer"])
    assert scaled_err(full, g["k_norm"][:nx, :ny]) < TOL


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
    assert scaled_err(gram, g["gram"][:nx, :nx]) < TOL
    sample = orc.random_phase_sample(space, meas, g["s_phases"], g["s_weights"], g["x"][:nx], g["es_normalizer"])
    assert scaled_err(sample, g["sample"][:nx]) < TOL
    if nx <= 12:
        kes = orc.eigensum(space, meas, g["x"][:6], g["y"][:6], g["es_normalizer"])
        assert scaled_err(kes, g["k_es"][:6, :6]) < TOL


@pytest.mark.parametrize("name", ["rpk_so3", "rpk_so5", "rpk_su3"])
def test_random_phase_on_groups(name):
    g = load_golden(name)
    space = orc.Group(g["group"], g["n"], g["order"])
    meas = _measure(g["measure"])
    emb = orc.random_phase_embedding(space, meas, g["phases"], g["x"], torch.sqrt)
    assert scaled_err(emb, g["embedding"]) < TOL
    gram = orc.random_phase_kernel(space, meas, g["phases"], g["x"], g["y"], g["normalizer"])
    assert scaled_err(gram, g["gram"]) < TOL
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
    assert scaled_err(gram, g["gram"]) < TOL
    sample = orc.rff_sample(space, meas, g["lmd"], g["shift"], g["w_re"], g["w_im"], g["x"], g["normalizer"])
    assert scaled_err(sample, g["sample"]) < TOL


@pytest.mark.parametrize("name", ["eigsum_torus2_matern_L20", "eigsum_torus3_heat_L31", "eigsum_torus1_heat_L15"])
def test_torus_eigensum_matches_reference(name):
    g = load_golden(name)
    space = orc.Torus(g["n"], g["order"])
    assert [(tuple(s), d, c) for s, d, c in space.irreps] == [(tuple(s), d, c) for s, d, c in g["irreps"]]
    meas = _measure(g["measure"])
    raw = orc.eigensum_unnormalised(space, meas, g["x"], g["y"])
    assert scaled_err(raw, g["k_raw"]) < TOL
    assert scaled_err(orc.eigensum(space, meas, g["x"], g["y"], g["normalizer"]), g["k_norm"]) < TOL
    assert torch.equal(space.pairwise_dist(g["x"][:16], g["y"][:12]), g["pairwise_dist"])
    assert torch.equal(space.dist(g["x"][:16], g["y"][:16]), g["dist"])
