"""The C-ABI library loads without a GPU and exports exactly the symbols include/lsk.h declares; the ctypes mirror of
`LskEpilogue` has the C layout; argument validation answers before any CUDA call."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "lsk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from liestationarykernels_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.exported_symbols()) == declared
    assert lib.lsk_version() >= 100


def test_epilogue_struct_layout_matches_header():
    from liestationarykernels_b200 import _lib
    # C layout: 4 + 4 + 4 + 24 + 4 int32 = 40 ints = 160 bytes, then 1 + 4*5 + 400 doubles
    assert ctypes.sizeof(_lib.LskEpilogue) == 160 + 8 * (2 + 20 + 400)
    assert _lib.LskEpilogue.scale.offset == 160
    assert _lib.LskEpilogue.coef.offset == 160 + 8 * 22
    assert ctypes.sizeof(_lib.LskEpilogue) < 4096, "must fit the kernel-parameter constant bank"


def test_argument_errors_without_gpu():
    from liestationarykernels_b200 import _lib
    lib = _lib.load()
    ep = _lib.LskEpilogue()
    assert lib.lsk_pairwise_poly(None, None, 4, 4, 1, ctypes.byref(ep), None, 4, None) == -1
    assert lib.lsk_embed_points(None, 4, 3, 0, 1, None, None, None, 3, None, None) == -1
    assert lib.lsk_lift_frames(None, 1, 5, 3, None, None, None) == -1
    assert lib.lsk_embed_buffer_doubles(1, 3) == 2 * 3 * 256
    assert lib.lsk_embed_buffer_doubles(129, 32) == 4 * 32 * 256


def test_compute_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from liestationarykernels_b200 import _lib
    from liestationarykernels_b200.spaces import SO
    with pytest.raises(_lib.LskError):
        SO(3, order=5)
