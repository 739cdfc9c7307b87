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


def test_wgrad_plan_host_logic(built_lib):
    """Weight-gradient tiling / operand orientation (csrc/wgrad.cu:wg_choose) at configs[1] shapes: groups of row tiles
    that share a dz stage are bounded by TMEM (G * bnt <= 512 columns) and by 3 pipeline stages of shared memory; the
    operands are exchanged only for stride-1 'same' convs where the MMA cost model says so."""
    import ctypes
    from selavi_b200 import ops

    def plan(nb, ci, co, thw, k, s, p):
        v = [ctypes.c_int() for _ in range(7)]
        assert built_lib.selavi_conv_wgrad_plan(ops.ConvGeom(nb, ci, co, thw, k, s, p).arr(0), ci, *[ctypes.byref(x) for x in v]) == 0
        return dict(zip(("mtiles", "bnt", "ntiles", "G", "slices", "exchanged", "ctas"), (x.value for x in v)))

    l1s = plan(16, 64, 144, (32, 56, 56), (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert (l1s["mtiles"], l1s["bnt"], l1s["ntiles"], l1s["G"], l1s["exchanged"]) == (5, 144, 1, 3, 0)
    # split-K: one wave of 148 CTAs, 74 slices for each of the two groups of row tiles (same pixels at the same time)
    assert (l1s["slices"], l1s["ctas"]) == (74, 148)
    l1t = plan(16, 144, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0))      # 4 row tiles x N=64 -> 2 row tiles x N=144
    assert (l1t["mtiles"], l1t["bnt"], l1t["G"], l1t["exchanged"]) == (2, 144, 2, 1)
    strided = plan(16, 64, 230, (32, 56, 56), (1, 3, 3), (1, 2, 2), (0, 1, 1))  # strided: never exchanged
    assert strided["exchanged"] == 0 and strided["bnt"] == 240
    for name, cfg in {"l2s": (16, 128, 288, (16, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                      "l4s": (16, 512, 1152, (4, 7, 7), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                      "l4t": (16, 1152, 512, (4, 7, 7), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                      "stem_t": (16, 45, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0))}.items():
        pl = plan(*cfg)
        assert 1 <= pl["G"] <= 3 and pl["G"] * pl["bnt"] <= 512, name
        assert pl["bnt"] % 16 == 0 and pl["bnt"] <= 256 and pl["slices"] >= 1, name
        groups = -(-pl["mtiles"] // pl["G"])
        assert groups * pl["ntiles"] <= pl["ctas"] <= max(148, groups * pl["ntiles"]) + groups, name   # one wave of CTAs


def test_halo_plan_host_logic(built_lib):
    """Tiling of the tap-reuse kernel (csrc/conv_halo.cu:hl_plan, host only): 128-pixel tiles (8 frames x 16 positions for
    3x1x1, 128 positions of the row-padded frame for 1x3x3), column tiles of <= 256 channels in multiples of 16, packed
    weights = tiles x chunks x taps x (hi|lo) x bnt x 128 B; strided / other kernels are refused (return 1)."""
    import ctypes
    from selavi_b200 import ops

    def plan(nb, ci, co, thw, k, s=(1, 1, 1), mode=0):
        p = (1, 0, 0) if k[0] == 3 else ((0, 1, 1) if k[1] == 3 else (0, 0, 0))
        mt, bnt, nt, wb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        code = built_lib.selavi_conv_halo_plan(ops.ConvGeom(nb, ci, co, thw, k, s, p).arr(mode), ctypes.byref(mt), ctypes.byref(bnt),
                                               ctypes.byref(nt), ctypes.byref(wb))
        return code, (mt.value, bnt.value, nt.value, wb.value)

    assert plan(16, 144, 64, (32, 56, 56), (3, 1, 1)) == (0, (16 * 4 * 196, 64, 1, 3 * 3 * 2 * 64 * 128))
    assert plan(16, 64, 144, (32, 56, 56), (1, 3, 3)) == (0, (16 * 32 * 26, 144, 1, 1 * 9 * 2 * 144 * 128))
    # 230 channels: one 240-wide column tile would leave room for a single 61 KB weight slot next to the two 48 KB windows
    # and the staging tiles, so the plan falls back to two 128-wide column tiles
    assert plan(16, 128, 230, (16, 28, 28), (1, 3, 3)) == (0, (1792, 128, 2, 2 * 2 * 9 * 2 * 128 * 128))
    assert plan(16, 128, 288, (16, 28, 28), (1, 3, 3)) == (0, (1792, 144, 2, 2 * 2 * 9 * 2 * 144 * 128))
    assert plan(16, 512, 1152, (4, 7, 7), (1, 3, 3)) == (0, (64, 192, 6, 6 * 8 * 9 * 2 * 192 * 128))
    # data-gradient mode of the same stride-1 convs: source / destination channels exchanged
    code, (mt, bnt, nt, _) = plan(16, 64, 144, (32, 56, 56), (1, 3, 3), mode=1)
    assert code == 0 and (mt, bnt, nt) == (16 * 32 * 26, 64, 1)
    for cfg in [(16, 64, 230, (32, 56, 56), (1, 3, 3), (1, 2, 2)), (16, 230, 128, (32, 28, 28), (3, 1, 1), (2, 1, 1)),
                (16, 64, 128, (32, 56, 56), (1, 1, 1), (2, 2, 2))]:
        assert plan(*cfg)[0] == 1
