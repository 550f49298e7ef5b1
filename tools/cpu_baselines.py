"""CPU baselines of every BASELINE config beside the GPU numbers: the oracle port of the reference
algorithm (oracle/lsk_oracle.py - test infrastructure, used here only as the thing timed on the host cores) on bounded
samples of each workload, all host threads.  Writes gpurun_out/cpu_baselines.json.
    python tools/cpu_baselines.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import lsk_oracle as orc

torch.set_num_threads(os.cpu_count() or 1)
res = {"cores": torch.get_num_threads(), "cpu_count": os.cpu_count(), "kind": "port (oracle/lsk_oracle.py)"}


def eig(tag, group, n, order, meas, nx, ny):
    torch.manual_seed(0)
    og = orc.Group(group, n, order)
    x, y = og.rand(nx), og.rand(ny)
    for irrep in og.irreps:                      # table construction outside the timed region
        og._table(irrep[0])
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.eigensum_unnormalised(og, meas, x, y)
    t = time.perf_counter() - t0
    res[tag] = {"sample": "%dx%d block, %d irreps" % (nx, ny, order), "seconds": t, "entries_per_s": nx * ny / t}


eig("C1_so3_heat_L20", "SO", 3, 20, orc.Measure("heat", 3, 1.0), 1024, 1024)
eig("C2_so5_matern_L100", "SO", 5, 100, orc.Measure("matern", 10, 1.0, 2.5), 32, 32)
eig("C2_so5_matern_L20", "SO", 5, 20, orc.Measure("matern", 10, 1.0, 2.5), 128, 128)
eig("C3_su4_heat_L20", "SU", 4, 20, orc.Measure("heat", 15, 1.0), 192, 192)

# C4: Stiefel V(5,3), feature map of N points against P phases (A = 30 subgroup samples, 10 irreps) and the sample
torch.manual_seed(0)
st = orc.Homogeneous("stiefel", 5, 3, order=10, average_order=30)
meas = orc.Measure("matern", 9, 1.0, 2.5)
N, P = 64, 32
x, phases = st.rand(N), st.rand(P)
for irrep in st.g.irreps:
    st.g._table(irrep[0])
t0 = time.perf_counter()
with torch.no_grad():
    emb = orc.random_phase_embedding(st, meas, phases, x, torch.sqrt)
t = time.perf_counter() - t0
res["C4_stiefel53_features"] = {"sample": "N=%d points x P=%d phases" % (N, P), "seconds": t, "point_phase_pairs_per_s": N * P / t}

# C5: spherical-function features on H^3 and SPD(3)
for tag, sp, meas in (("C5_h3_features", orc.Hyperbolic(3, order=8192), orc.Measure("matern", 3, 1.0, 2.5)),
                      ("C5_spd3_features", orc.SPD(3, order=8192), orc.Measure("matern", 6, 1.0, 2.5))):
    torch.manual_seed(0)
    N = 2048 if tag == "C5_h3_features" else 16
    x = sp.rand(N)
    lmd, shift = sp.generate(meas)
    t0 = time.perf_counter()
    with torch.no_grad():
        f = sp.features(sp.to_group(x), lmd, shift)
    t = time.perf_counter() - t0
    res[tag] = {"sample": "N=%d points x m=8192 features" % N, "seconds": t, "features_per_s": N * 8192 / t}
    # Gram of the feature maps: torch CPU DGEMM on a bounded block
    psi = torch.cat((f.real, f.imag), dim=-1)
    t0 = time.perf_counter()
    g = psi @ psi.t()
    t = time.perf_counter() - t0
    res[tag]["gram_block_tflops"] = 2.0 * N * N * psi.shape[1] / t / 1e12

print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "cpu_baselines.json"), "w"), indent=1)
