"""Development: one C4 feature map (V(5,3), N = 32768, P = 2048, A = 30) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.spaces import Stiefel
from liestationarykernels_b200.spectral_kernel import RandomPhaseKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
torch.manual_seed(0)
space = Stiefel(5, 3)
kern = RandomPhaseKernel(MaternSpectralMeasure(space.dim, 1.0, 2.5), space, phase_order=2048).eval()
x = space.rand(32768)
from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
sampler = RandomPhaseApproximation(EigenbasisSumKernel(MaternSpectralMeasure(space.dim, 1.0, 2.5), space).eval(), phase_order=2048)
with torch.no_grad():
    f = kern._features(x, kern._row_scales(), panel=True)
    s = sampler(x)
torch.cuda.synchronize()
print("ok", f.rows, f.cols, s.shape)
