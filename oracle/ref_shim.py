"""TEST INFRASTRUCTURE — import the *real* reference (imbirik/LieStationaryKernels) in this container.

The reference is pure Python but pulls three packages that are not in the image
(`gpytorch.lazy`, `geomstats`, `more_itertools`, `spherical_harmonics`).  None of them contributes
arithmetic to the hot path, so tiny stand-ins registered in `sys.modules` are enough to
import it unmodified.  The packaged character table holds 20 irreps per group; the 100-irrep SO(3..5)
table ships under `plots/`.  The reference tree is read-only, so the package is copied to a scratch
directory and the two JSON tables are merged there — the same thing the reference's own notebooks do
(`plots/so_figure_8_1.ipynb`, cell 3).

Only `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py` (skipped when `/root/reference`
is absent, e.g. on the GPU box) use this module.  Nothing under `liestationarykernels_b200/` may.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import types
import warnings

REFERENCE_ROOT = os.environ.get("LSK_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "lie_stationary_kernels"))


def _install_stubs() -> None:
    import torch

    if "gpytorch" not in sys.modules:
        gpytorch = types.ModuleType("gpytorch")
        lazy = types.ModuleType("gpytorch.lazy")

        class NonLazyTensor:
            def __init__(self, tensor):
                self.tensor = tensor

            def t(self):
                return NonLazyTensor(self.tensor.t())

            def clone(self):
                return NonLazyTensor(self.tensor.clone())

            def evaluate(self):
                return self.tensor

            @property
            def shape(self):
                return self.tensor.shape

        class MatmulLazyTensor:
            def __init__(self, left, right):
                self.left, self.right = left, right

            def evaluate(self):
                return self.left.tensor @ self.right.tensor

            def __getitem__(self, idx):
                return self.evaluate()[idx]

            @property
            def shape(self):
                return torch.Size((self.left.tensor.shape[0], self.right.tensor.shape[1]))

        lazy.NonLazyTensor = NonLazyTensor
        lazy.MatmulLazyTensor = MatmulLazyTensor
        gpytorch.lazy = lazy
        sys.modules["gpytorch"] = gpytorch
        sys.modules["gpytorch.lazy"] = lazy

    if "geomstats" not in sys.modules:
        geomstats = types.ModuleType("geomstats")
        geometry = types.ModuleType("geomstats.geometry")
        sys.modules["geomstats"] = geomstats
        sys.modules["geomstats.geometry"] = geometry
        for mod, cls in (("stiefel", "Stiefel"), ("grassmannian", "Grassmannian"), ("hypersphere", "Hypersphere")):
            m = types.ModuleType("geomstats.geometry." + mod)

            class _Base:
                def __init__(self, *a, **k):
                    pass

            _Base.__name__ = cls
            setattr(m, cls, _Base)
            sys.modules["geomstats.geometry." + mod] = m
            setattr(geometry, mod, m)
        geomstats.geometry = geometry

    if "more_itertools" not in sys.modules:
        mi = types.ModuleType("more_itertools")

        def always_iterable(obj):
            try:
                return iter(obj)
            except TypeError:
                return iter((obj,))

        def distinct_permutations(it):
            import itertools
            return iter(sorted(set(itertools.permutations(it))))

        mi.always_iterable = always_iterable
        mi.distinct_permutations = distinct_permutations
        sys.modules["more_itertools"] = mi

    if "spherical_harmonics" not in sys.modules:
        sh = types.ModuleType("spherical_harmonics")
        shs = types.ModuleType("spherical_harmonics.spherical_harmonics")
        fs = types.ModuleType("spherical_harmonics.fundamental_set")

        class _Dummy:
            def __init__(self, *a, **k):
                pass

        def num_harmonics(dimension, degree):
            """Published formula of the absent dependency (vdutor/SphericalHarmonics, `num_harmonics`): the number of
            spherical harmonics of a degree in R^dimension.  It only feeds `eigenspace.dimension` of the reference's
            Sphere (spaces/sphere.py:122-123); the kernel values do not depend on it."""
            import math
            if degree == 0:
                return 1
            if dimension == 3:
                return int(2 * degree + 1)
            return int(round((2 * degree + dimension - 2) / degree * math.comb(degree + dimension - 3, degree - 1)))

        shs.SphericalHarmonicsLevel = _Dummy
        shs.num_harmonics = num_harmonics
        fs.FundamentalSystemCache = _Dummy
        sh.spherical_harmonics = shs
        sh.fundamental_set = fs
        sys.modules["spherical_harmonics"] = sh
        sys.modules["spherical_harmonics.spherical_harmonics"] = shs
        sys.modules["spherical_harmonics.fundamental_set"] = fs


_SCRATCH = None


def import_reference():
    """Returns the imported `lie_stationary_kernels` package of the reference (CPU torch)."""
    global _SCRATCH
    if "lie_stationary_kernels" in sys.modules and getattr(sys.modules["lie_stationary_kernels"], "_lsk_is_reference", False):
        return sys.modules["lie_stationary_kernels"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    _SCRATCH = tempfile.mkdtemp(prefix="lsk_ref_")
    src = os.path.join(REFERENCE_ROOT, "src", "lie_stationary_kernels")
    dst = os.path.join(_SCRATCH, "lie_stationary_kernels")
    shutil.copytree(src, dst)
    packaged = os.path.join(dst, "spaces", "precomputed_characters.json")
    with open(packaged) as fh:
        table = json.load(fh)
    with open(os.path.join(REFERENCE_ROOT, "plots", "precomputed_characters.json")) as fh:
        big = json.load(fh)
    for group, irreps in big.items():
        table.setdefault(group, {})
        for sig, val in irreps.items():
            if sig in table[group]:
                assert table[group][sig] == val, (group, sig)
            table[group][sig] = val
    with open(packaged, "w") as fh:
        json.dump(table, fh)
    # drop any non-reference module of the same name (the drop-in alias package of this repo)
    for name in [m for m in sys.modules if m == "lie_stationary_kernels" or m.startswith("lie_stationary_kernels.")]:
        del sys.modules[name]
    sys.path.insert(0, _SCRATCH)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import lie_stationary_kernels  # noqa: F401
            import lie_stationary_kernels.spaces  # noqa: F401
            import lie_stationary_kernels.spectral_kernel  # noqa: F401
            import lie_stationary_kernels.spectral_measure  # noqa: F401
            import lie_stationary_kernels.prior_approximation  # noqa: F401
    finally:
        sys.path.remove(_SCRATCH)
    pkg = sys.modules["lie_stationary_kernels"]
    pkg._lsk_is_reference = True
    return pkg


def merged_reference_table() -> dict:
    """The merged (packaged ∪ plots) reference character table, for validating our own generator."""
    with open(os.path.join(REFERENCE_ROOT, "src", "lie_stationary_kernels", "spaces", "precomputed_characters.json")) as fh:
        table = json.load(fh)
    with open(os.path.join(REFERENCE_ROOT, "plots", "precomputed_characters.json")) as fh:
        big = json.load(fh)
    for group, irreps in big.items():
        table.setdefault(group, {}).update(irreps)
    return table
