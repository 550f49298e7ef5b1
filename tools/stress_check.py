"""Development: determinism / self-consistency stress of the pair kernels (races in the claimed-producer ring or the
whole-tile pipeline would show up as run-to-run differences or as block-vs-full mismatches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liestationarykernels_b200.spaces import SO, SU, Torus
from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure

torch.manual_seed(0)
bad = 0
cases = [(SO(5, order=100), MaternSpectralMeasure(10, 1.0, 2.5), 4099, 3001), (SO(5, order=20), SqExpSpectralMeasure(10, 1.0), 5000, 777),
         (SO(5, order=37), MaternSpectralMeasure(10, 1.0, 2.5), 3000, 2100), (SU(4, order=20), SqExpSpectralMeasure(15, 1.0), 3001, 2049),
         (SU(3, order=13), SqExpSpectralMeasure(8, 1.0), 2500, 3100), (SO(7, order=20), SqExpSpectralMeasure(21, 1.0), 700, 900),
         (SO(3, order=50), SqExpSpectralMeasure(3, 0.5), 6000, 4097), (SU(6, order=10), SqExpSpectralMeasure(35, 1.0), 500, 300),
         (Torus(2, order=40), MaternSpectralMeasure(2, 0.3, 2.5), 5000, 3000)]
with torch.no_grad():
    for space, meas, nx, ny in cases:
        kern = EigenbasisSumKernel(meas, space).eval()
        x, y = space.rand(nx), space.rand(ny)
        ref = kern(x, y)
        for it in range(8):
            again = kern(x, y)
            if not torch.equal(again, ref):
                bad += 1
                print("NON-DETERMINISTIC", type(space).__name__, space.n, it, (again - ref).abs().max().item())
        # row/column blocks must reproduce the full matrix exactly (same per-entry arithmetic, different tiling)
        r0, r1, c0, c1 = nx // 3 + 1, nx // 3 + 517, ny // 5 + 3, ny // 5 + 300
        blk = kern(x[r0:r1].contiguous(), y[c0:c1].contiguous())
        if not torch.equal(blk, ref[r0:r1, c0:c1]):
            d = (blk - ref[r0:r1, c0:c1]).abs().max().item()
            print("BLOCK MISMATCH", type(space).__name__, space.n, d)
            bad += d > 1e-13 * ref.abs().max().item()
        sym = kern(x[:2500])
        gen = kern(x[:2500], x[:2500].clone())
        if not torch.equal(sym, gen):
            print("SYM MISMATCH", type(space).__name__, space.n, (sym - gen).abs().max().item())
            bad += 1
        print("ok", type(space).__name__, space.n, len(space.lb_eigenspaces), flush=True)
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
