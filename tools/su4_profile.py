"""Development: one C3 covariance (SU(4) heat, 16384^2, 20 irreps) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.spaces import SU
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
torch.manual_seed(0)
sp = SU(4, order=20)
k = EigenbasisSumKernel(SqExpSpectralMeasure(15, 1.0), sp).eval()
x, y = sp.rand(16384), sp.rand(16384)
with torch.no_grad():
    out = k(x, y)
torch.cuda.synchronize()
print("ok", out.shape)
