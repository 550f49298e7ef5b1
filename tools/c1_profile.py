"""Development: C1 covariance calls (SO(3) heat, 1024^2, 20 irreps) for launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.spaces import SO
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import SqExpSpectralMeasure
torch.manual_seed(0)
sp = SO(3, order=20)
k = EigenbasisSumKernel(SqExpSpectralMeasure(3, 1.0), sp).eval()
x, y = sp.rand(1024), sp.rand(1024)
with torch.no_grad():
    for _ in range(3):
        out = k(x, y)
torch.cuda.synchronize()
print("ok", out.shape)
