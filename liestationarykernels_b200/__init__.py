"""B200-native kernel evaluation and prior sampling for stationary kernels on Lie groups and symmetric spaces.

Drop-in for the hot path of imbirik/LieStationaryKernels: same Python classes, computed by hand-written sm_100a
CUDA kernels behind a C ABI (liblsk.so).  No CPU fallback."""
from . import _lib  # noqa: F401
