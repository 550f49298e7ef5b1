"""Drop-in import name of the reference package (`lie_stationary_kernels`), backed by the B200-native implementation
in `liestationarykernels_b200`.  `from lie_stationary_kernels.spaces import SO` etc. work unchanged."""
import sys as _sys

import liestationarykernels_b200 as _impl
from liestationarykernels_b200 import prior_approximation, space, spaces, spectral_kernel, spectral_measure, utils  # noqa: F401

for _name in ("prior_approximation", "space", "spaces", "spectral_kernel", "spectral_measure", "utils"):
    _sys.modules[__name__ + "." + _name] = getattr(_impl, _name) if hasattr(_impl, _name) else _sys.modules["liestationarykernels_b200." + _name]
for _sub in ("so", "su", "stiefel", "grassmannian", "hyperbolic", "spd", "torus", "sphere"):
    import importlib as _il
    _sys.modules[__name__ + ".spaces." + _sub] = _il.import_module("liestationarykernels_b200.spaces." + _sub)
