"""Development: one C2 covariance (SO(5) Matern-5/2, 16384^2, 100 irreps) for ncu captures of so5::stair_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
torch.manual_seed(0)
sp = SO(5, order=100)
k = EigenbasisSumKernel(MaternSpectralMeasure(10, 1.0, 2.5), sp).eval()
x, y = sp.rand(16384), sp.rand(16384)
with torch.no_grad():
    out = k(x, y)
torch.cuda.synchronize()
print("ok", out.shape)
