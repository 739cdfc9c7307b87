"""CPU checks of the drop-in boundary: the library builds, loads, and exports exactly what the header declares."""
import os
import re

from selavi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "selavi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(selavi_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    names = _header_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/selavi_b200.h but not exported"


def test_ctypes_signatures_cover_header(built_lib):
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_version_and_error_string(built_lib):
    assert built_lib.selavi_version() >= 100
    assert isinstance(built_lib.selavi_last_error(), bytes)


def test_host_only_queries(built_lib):
    # pure host helpers: tile selection and workspace sizes (no device needed)
    import ctypes
    bnt, nt = ctypes.c_int(), ctypes.c_int()
    for n_out, exp in [(45, (48, 1)), (64, (64, 1)), (144, (144, 1)), (230, (240, 1)), (288, (144, 2)),
                       (460, (240, 2)), (576, (192, 3)), (921, (240, 4)), (1152, (240, 5)), (512, (256, 2))]:
        assert built_lib.selavi_conv_tiles(n_out, ctypes.byref(bnt), ctypes.byref(nt)) == 0
        assert (bnt.value, nt.value) == exp
        assert bnt.value % 16 == 0 and bnt.value <= 256 and bnt.value * nt.value >= n_out
    assert built_lib.selavi_sk_kp(309) == 352 and built_lib.selavi_sk_kp(28) == 64 and built_lib.selavi_sk_kp(513) == -1
    assert built_lib.selavi_sk_workspace_bytes(309) > 0
    assert built_lib.selavi_conv_wpack_bytes(144, 576) == 18 * 2 * 144 * 128


def test_argument_validation_without_gpu(built_lib):
    # argument errors are reported before any CUDA call, with a message
    code = built_lib.selavi_sk_solve(None, 0, 0, 309, 20.0, 0, None, None, None, None, None, 10, 10, 0.1, 1, 1, 1, None,
                                     None, None, 1, 0, None, None, None)
    assert code < 0 and b"sk" in built_lib.selavi_last_error()
