"""Development: builds liblsk_trace.so (pairwise kernel with -DLSK_TRACE clock stamps) next to the product library."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "liestationarykernels_b200", "csrc")
out = os.path.join(ROOT, "tools", "liblsk_trace.so")
srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
       "--expt-relaxed-constexpr", "-DLSK_TRACE",] + sys.argv[1:] + [ "-I", os.path.join(ROOT, "include"), "-shared", "-o", out] + srcs + ["-lcudart"]
subprocess.check_call(cmd)
print(out)
