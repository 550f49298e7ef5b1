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


def test_epilogue_struct_layout_matches_header(tmp_path):
    """The ctypes mirror of LskEpilogue against what a C compiler makes of include/lsk.h."""
    import os
    import subprocess
    from liestationarykernels_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lsk.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(LskEpilogue), offsetof(LskEpilogue, group_end), offsetof(LskEpilogue, row_len), offsetof(LskEpilogue, scale), '
                   'offsetof(LskEpilogue, aff), offsetof(LskEpilogue, coef)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    c = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    E = _lib.LskEpilogue
    assert c == [ctypes.sizeof(E), E.group_end.offset, E.row_len.offset, E.scale.offset, E.aff.offset, E.coef.offset]
    # 4 + 5 + 5 + 24 + 4 int32 = 168 bytes, then scale, stair_shape, 5 x 6 affine map, 400 coefficients
    assert E.scale.offset == 168 and E.coef.offset == 168 + 8 * 32
    assert ctypes.sizeof(E) < 4096, "must fit the kernel-parameter constant bank"


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


def test_error_strings_are_stateless_and_cover_the_codes():
    """lsk_error_string needs no device: the text of every LSK_E* code and of a CUDA error code"""
    from liestationarykernels_b200 import _lib
    lib = _lib.load()
    assert lib.lsk_error_string(0) == b"ok"
    assert lib.lsk_error_string(-1) == b"invalid argument"
    assert lib.lsk_error_string(-2) == b"size out of supported range"
    assert lib.lsk_error_string(-3) == b"configuration not compiled"
    assert lib.lsk_error_string(-99) == b"unknown error code"
    assert lib.lsk_error_string(2)            # cudaErrorMemoryAllocation: some non-empty text from the runtime
    with pytest.raises(_lib.LskError, match="size out of supported range"):
        _lib.check(-2, "probe")
    assert lib.lsk_potrf_work_doubles(0) >= 0 and lib.lsk_potrf_work_doubles(300) > lib.lsk_potrf_work_doubles(0)
    assert lib.lsk_potri_work_doubles(-1) == -2 and lib.lsk_trsm_work_doubles(128, 5) > 0


def test_python_embed_size_matches_the_library():
    """ops._embed_doubles mirrors lsk_embed_buffer_doubles (the hot call computes it without a library round trip)"""
    from liestationarykernels_b200 import _lib, ops
    lib = _lib.load()
    for num in (1, 63, 64, 127, 128, 129, 1000, 16384, 65536):
        for kg in (1, 3, 11, 21, 4096):
            assert ops._embed_doubles(num, kg) == lib.lsk_embed_buffer_doubles(num, kg)
