"""GP regression on the manifolds of the reference's `examples/gpr_experiments.py`, on this package (no gpytorch):
a prior sample as ground truth, 30 training points, hyper-parameters fitted by Adam on the exact marginal likelihood,
mean squared error on 100 test points against the variance of the targets.

    python examples/gpr_experiments.py [--iters 300]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from liestationarykernels_b200 import gpr
from liestationarykernels_b200.prior_approximation import RandomFourierApproximation, RandomPhaseApproximation
from liestationarykernels_b200.space import CompactLieGroup
from liestationarykernels_b200.spaces import (SO, SU, Grassmannian, HyperbolicSpace, ProjectiveSpace, Sphere, Stiefel,
                                              SymmetricPositiveDefiniteMatrices)
from liestationarykernels_b200.spaces.homogeneous import CompactHomogeneousSpace
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomFourierFeatureKernel, RandomPhaseKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure


def make_kernel(space, measure):
    if isinstance(space, (CompactLieGroup, Sphere)):
        return EigenbasisSumKernel(measure, space)
    if isinstance(space, CompactHomogeneousSpace):
        return RandomPhaseKernel(measure, space, phase_order=24)
    return RandomFourierFeatureKernel(measure, space)


def main(space, shape, iters):
    truth_kernel = make_kernel(space, MaternSpectralMeasure(space.dim, 0.5, 1.5, 4.0))
    if isinstance(truth_kernel, RandomFourierFeatureKernel):
        f = RandomFourierApproximation(truth_kernel)
    else:
        f = RandomPhaseApproximation(truth_kernel)
    n_train, n_test = 30, 100
    train_x, test_x = space.rand(n_train), space.rand(n_test)
    with torch.no_grad():
        truth_kernel.eval()
        train_y, test_y = f(train_x), f(test_x)
    train_x, test_x = train_x.reshape(n_train, -1), test_x.reshape(n_test, -1)
    kernel = make_kernel(space, MaternSpectralMeasure(space.dim, 1.0, 1.5))
    likelihood = gpr.GaussianLikelihood(device="cuda")
    model = gpr.ExactGPModel(train_x, train_y, likelihood, kernel, space, point_shape=shape)
    losses = gpr.train(model, train_x, train_y, iters, max(iters // 2, 1), lr=0.1, verbose=False)
    model.eval()
    kernel.eval()
    with torch.no_grad():
        pred = model(test_x)
    mse = ((pred.mean - test_y) ** 2).mean().item()
    var = test_y.var(unbiased=False).item()
    mll = gpr.ExactMarginalLogLikelihood(likelihood, model)(pred, test_y).item()
    print("%-38s loss %7.3f -> %7.3f   lengthscale %.3f  noise %.4f   test mse %.4f  (target variance %.4f)  test mll %.3f" % (
        "%s%s" % (type(space).__name__, tuple(shape) if shape else (space.n,)), losses[0], losses[-1],
        kernel.measure.lengthscale.item(), likelihood.noise.item(), mse, var, mll), flush=True)
    return mse, var


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    torch.manual_seed(0)
    spaces = [(SU(2), (2, 2)), (SU(3), (3, 3)), (SO(3), (3, 3)), (SO(4), (4, 4)), (SO(5), (5, 5)),
              (Stiefel(5, 3), (5, 3)), (Grassmannian(3, 1), (3, 1)), (Grassmannian(4, 2), (4, 2)),
              (Sphere(2), None), (Sphere(3), None), (ProjectiveSpace(2), None),
              (HyperbolicSpace(2, order=2000), None), (HyperbolicSpace(3, order=2000), None),
              (SymmetricPositiveDefiniteMatrices(2), (2, 2)), (SymmetricPositiveDefiniteMatrices(3), (3, 3))]
    for space, shape in spaces:
        main(space, shape, args.iters)
