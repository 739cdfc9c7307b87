"""Thin checked Python wrappers over the C-ABI kernels (one function per entry point of include/selavi_b200.h).

Tensors are torch CUDA tensors used purely as device memory; every function launches on the current stream
and returns without synchronising.  Activations are channels-last fp32 `[N, T, H, W, Cs]` with `Cs` = channel
count padded to a multiple of 8 (pad channels are zero).  No CPU fallback: a missing library raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class _Guard:
    """cheap device guard: only switches (and restores) the current device when the tensor lives elsewhere"""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx, self.prev = dev.index, None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# bench.py sets PROFILE to a list: every conv launch is then bracketed by CUDA events on the launching stream
PROFILE = None


class _Prof:
    def __init__(self, kind, geom, kernel):
        self.on = PROFILE is not None
        if self.on:
            self.kernel = kernel
            self.kind, self.flops = kind, 2.0 * geom.m_out * geom.co * geom.ci * geom.taps
            self.tag = (geom.ci, geom.co, geom.ti, geom.hi, geom.wi, geom.kt, geom.kh, geom.kw, geom.st, geom.sh)
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        if self.on:
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            PROFILE.append((self.kind, self.flops, self.e0, self.e1, self.tag, self.kernel))


def padc(c):
    """channel stride of an activation with c channels: padded to 8 (32-byte rows of 8 fp32 / 16-byte units of 8 bf16)"""
    return (c + 7) & ~7


pad4 = padc  # historical name


def to_channels_last(x):
    """[N,C,T,H,W] or [N,C,H,W] (any float dtype) -> contiguous fp32 [N,T,H,W,padc(C)]."""
    if x.dim() == 4:
        x = x.unsqueeze(2)
    n, c, t, h, w = x.shape
    out = torch.zeros((n, t, h, w, padc(c)), dtype=torch.float32, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 4, 1)
    return out


def from_channels_last(x_cl, c):
    """[N,T,H,W,Cs] -> [N,C,T,H,W]."""
    return x_cl[..., :c].permute(0, 4, 1, 2, 3).contiguous()


class ConvGeom:
    """Forward geometry of one convolution (3-D; 2-D convs use T=1, kt=1)."""

    def __init__(self, nb, ci, co, in_thw, kernel, stride, padding):
        self.nb, self.ci, self.co = nb, ci, co
        self.ti, self.hi, self.wi = in_thw
        self.kt, self.kh, self.kw = kernel
        self.st, self.sh, self.sw = stride
        self.pt, self.ph, self.pw = padding
        self.to = (self.ti + 2 * self.pt - self.kt) // self.st + 1
        self.ho = (self.hi + 2 * self.ph - self.kh) // self.sh + 1
        self.wo = (self.wi + 2 * self.pw - self.kw) // self.sw + 1
        self.cis, self.cos = padc(ci), padc(co)
        self.taps = self.kt * self.kh * self.kw
        self.m_out = nb * self.to * self.ho * self.wo
        self.m_in = nb * self.ti * self.hi * self.wi

    def arr(self, mode):
        if mode == 0:
            g = [0, self.nb, self.ti, self.hi, self.wi, self.cis, self.to, self.ho, self.wo, self.cos,
                 self.kt, self.kh, self.kw, self.st, self.sh, self.sw, self.pt, self.ph, self.pw, self.co]
        else:
            g = [1, self.nb, self.to, self.ho, self.wo, self.cos, self.ti, self.hi, self.wi, self.cis,
                 self.kt, self.kh, self.kw, self.st, self.sh, self.sw, self.pt, self.ph, self.pw, self.ci]
        return (ctypes.c_int * 20)(*g)

    def out_shape(self):
        return (self.nb, self.to, self.ho, self.wo, self.cos)

    def in_shape(self):
        return (self.nb, self.ti, self.hi, self.wi, self.cis)


def conv_tiles(n_out):
    bnt, nt = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().selavi_conv_tiles(n_out, ctypes.byref(bnt), ctypes.byref(nt)), "selavi_conv_tiles")
    return bnt.value, nt.value


def pack_weights(w, geom, mode, out=None):
    """w: torch weight [co, ci, kt, kh, kw] (or [co, ci, kh, kw]) fp32 CUDA -> packed B operand (uint8 buffer)."""
    w = w.detach()
    if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
        raise ValueError("weight must be a contiguous fp32 CUDA tensor")
    lib = _lib.lib()
    n_out, cs = (geom.ci, geom.cos) if mode == 1 else (geom.co, geom.cis)
    # mode 0 / 1: tf32 hi/lo tiles of the forward / data-gradient implicit GEMM; mode 2: fp16 hi/lo tiles of the forward
    nbytes = lib.selavi_conv_wpack_bytes_f16(n_out, geom.taps * cs) if mode == 2 else lib.selavi_conv_wpack_bytes(n_out, geom.taps * cs)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    elif out.numel() != nbytes:
        raise ValueError("packed weight buffer has the wrong size")
    with _Guard(w.device):
        _lib.check(lib.selavi_conv_pack_weights(_lib.ptr(w), mode, geom.co, geom.ci, geom.taps, cs, _lib.ptr(out),
                                                _lib.stream_ptr()), "selavi_conv_pack_weights")
    return out


# Forward kernel selection: "auto" = tap-reuse fp16x3 kernel (conv_halo.cu) where it is supported and pays
# (stride-1 3x1x1 / 1x3x3 convs with enough frames / wide enough rows), tf32x3 implicit GEMM (conv.cu) elsewhere;
# "igemm" = conv.cu everywhere; "halo" = conv_halo.cu wherever the geometry is supported.
import os as _os
FWD_KERNEL = _os.environ.get("SELAVI_FWD_KERNEL", "auto")
HALO_FLAGS = int(_os.environ.get("SELAVI_HALO_FLAGS", "0"))


def halo_plan(geom, mode=0):
    """-> (m_tiles, bnt, ntiles, wpack_bytes) of the tap-reuse kernel (mode 0 forward, 1 data gradient), or None when
    it does not apply."""
    key = "_halo%d" % mode
    cached = geom.__dict__.get(key)
    if cached is not None:
        return cached or None
    plan = False
    if FWD_KERNEL != "igemm":
        mt, bnt, nt, wb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        code = _lib.lib().selavi_conv_halo_plan(geom.arr(mode), ctypes.byref(mt), ctypes.byref(bnt), ctypes.byref(nt),
                                                ctypes.byref(wb))
        if code < 0:
            _lib.check(code, "selavi_conv_halo_plan")
        if code == 0:
            pays = True
            if FWD_KERNEL == "auto":
                # measured on B200 (tools/quick_bench.py halo, profiles/r01g_halo_microbench.txt): ~2x over the tf32x3
                # implicit GEMM on every R(2+1)D-18 layer except the 7x7 spatial convs of layer4, where only W/(W+2)
                # = 78 % of the MMA rows and 49/63 of each frame's last tile are useful
                if mode == 0:
                    pays = geom.kt == 3 or geom.wi >= 12
                else:
                    # data gradient (cp.async-fed in both kernels): the tap-reuse kernel wins on the temporal convs with
                    # at least one full 8-frame block (1.3-2.1x) and, since the coalesced accumulator drain, on every
                    # spatial conv (layer-1 64<-144: 1.06 ms vs 1.7 ms, gpurun r01k)
                    pays = geom.ti >= 8 if geom.kt == 3 else True
            if pays:
                plan = (mt.value, bnt.value, nt.value, wb.value)
    geom.__dict__[key] = plan
    return plan or None


def pack_weights_halo(w, geom, out=None, mode=0):
    """torch weight -> per-tap hi/lo B tiles of the tap-reuse kernel (mode 0: fp16, forward; mode 1: bf16, data gradient)."""
    w = w.detach()
    if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
        raise ValueError("weight must be a contiguous fp32 CUDA tensor")
    plan = halo_plan(geom, mode)
    if plan is None:
        raise ValueError("geometry not supported by the halo kernel")
    if out is None:
        out = torch.empty(plan[3], dtype=torch.uint8, device=w.device)
    elif out.numel() != plan[3]:
        raise ValueError("packed weight buffer has the wrong size")
    with _Guard(w.device):
        _lib.check(_lib.lib().selavi_conv_halo_pack_weights(_lib.ptr(w), geom.arr(mode), geom.ci if mode == 0 else geom.co,
                                                            _lib.ptr(out), _lib.stream_ptr()), "selavi_conv_halo_pack_weights")
    return out


def conv_dgrad_halo(z_hi, z_lo, wpack, geom, out=None, accumulate=False, bn=None):
    """bn = (z, scale, shift, mean, invstd) of the unit that produced the activation whose gradient this is: the epilogue
    then also emits that unit's BatchNorm-backward partial sums; returns (dx, stats_partial [m_tiles, 2, bnt*ntiles])."""
    _chk_bf16(z_hi, geom.out_shape(), "z_hi")
    _chk_bf16(z_lo, geom.out_shape(), "z_lo")
    if out is None:
        out = torch.empty(geom.in_shape(), dtype=torch.float32, device=z_hi.device)
        accumulate = False
    _chk(out, geom.in_shape(), "dx")
    if bn is not None:
        if accumulate:
            raise ValueError("fused BatchNorm-backward statistics need a written (not accumulated) gradient")
        bz, bsc, bsh, bmu, bis = bn
        _chk(bz, geom.in_shape(), "bn z")
        mt, bnt, nt, _ = halo_plan(geom, 1)
        stats = torch.empty((mt, 2, bnt * nt), dtype=torch.float32, device=z_hi.device)
        with _Guard(z_hi.device), _Prof("conv_dgrad", geom, "conv_halo_kernel[dgrad bf16x3]"):
            _lib.check(_lib.lib().selavi_conv_halo_dgrad_bnstats(_lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(out), _lib.ptr(wpack),
                                                                 geom.arr(1), _lib.ptr(bz), _lib.ptr(bsc), _lib.ptr(bsh), _lib.ptr(bmu),
                                                                 _lib.ptr(bis), _lib.ptr(stats), HALO_FLAGS, _lib.stream_ptr()),
                       "selavi_conv_halo_dgrad_bnstats")
        return out, stats
    with _Guard(z_hi.device), _Prof("conv_dgrad", geom, "conv_halo_kernel[dgrad bf16x3]"):
        _lib.check(_lib.lib().selavi_conv_halo_dgrad(_lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(out), _lib.ptr(wpack), geom.arr(1),
                                                     1 if accumulate else 0, HALO_FLAGS, _lib.stream_ptr()),
                   "selavi_conv_halo_dgrad")
    return out


def conv_forward_halo(x, wpack, geom, out=None, scale=None, shift=None, relu=False, stats=None):
    _chk(x, geom.in_shape(), "x")
    if out is None:
        out = torch.empty(geom.out_shape(), dtype=torch.float32, device=x.device)
    _chk(out, geom.out_shape(), "out")
    with _Guard(x.device), _Prof("conv_fwd", geom, "conv_halo_kernel[fwd fp16x3]"):
        _lib.check(_lib.lib().selavi_conv_halo_fwd(_lib.ptr(x), _lib.ptr(out), _lib.ptr(wpack), geom.arr(0), _lib.ptr(scale),
                                                   _lib.ptr(shift), 1 if relu else 0, _lib.ptr(stats), HALO_FLAGS,
                                                   _lib.stream_ptr()), "selavi_conv_halo_fwd")
    return out


def stats_buffer(geom, device, halo=False):
    if halo:
        mt, bnt, nt, _ = halo_plan(geom)
        return torch.empty((mt, 2, bnt * nt), dtype=torch.float32, device=device)
    bnt, nt = conv_tiles(geom.co)
    return torch.empty(((geom.m_out + 127) // 128, 2, bnt * nt), dtype=torch.float32, device=device)


def _chk(t, shape, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
        raise ValueError(f"{name}: expected contiguous fp32 CUDA tensor of shape {tuple(shape)}, got {tuple(t.shape)} {t.dtype}")


def conv_forward(x, wpack, geom, out=None, scale=None, shift=None, relu=False, stats=None, passes=3):
    _chk(x, geom.in_shape(), "x")
    if out is None:
        out = torch.empty(geom.out_shape(), dtype=torch.float32, device=x.device)
    _chk(out, geom.out_shape(), "out")
    with _Guard(x.device), _Prof("conv_fwd", geom, "conv_igemm_kernel[fwd fp16x3]" if passes == 6 else "conv_igemm_kernel[fwd tf32x3]"):
        _lib.check(_lib.lib().selavi_conv_gemm(_lib.ptr(x), _lib.ptr(out), _lib.ptr(wpack), geom.arr(0), _lib.ptr(scale),
                                               _lib.ptr(shift), 1 if relu else 0, _lib.ptr(stats), 0, passes,
                                               _lib.stream_ptr()), "selavi_conv_gemm(fwd)")
    return out


def conv_dgrad(dz, wpack_t, geom, out=None, accumulate=False, passes=3):
    _chk(dz, geom.out_shape(), "dz")
    if out is None:
        out = torch.empty(geom.in_shape(), dtype=torch.float32, device=dz.device)
        accumulate = False
    _chk(out, geom.in_shape(), "dx")
    with _Guard(dz.device), _Prof("conv_dgrad", geom, "conv_igemm_kernel[dgrad tf32x3]"):
        _lib.check(_lib.lib().selavi_conv_gemm(_lib.ptr(dz), _lib.ptr(out), _lib.ptr(wpack_t), geom.arr(1), None, None, 0,
                                               None, 1 if accumulate else 0, passes, _lib.stream_ptr()),
                   "selavi_conv_gemm(dgrad)")
    return out


_wgrad_ws = {}


def conv_wgrad(x, dz, geom, dw, scale=None, shift=None, relu=False, accumulate=False, passes=3):
    """dw: torch-layout gradient buffer [co, ci, kt, kh, kw] fp32, written (or accumulated into)."""
    _chk(x, geom.in_shape(), "x")
    _chk(dz, geom.out_shape(), "dz")
    if not (dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == geom.co * geom.ci * geom.taps):
        raise ValueError("dw must be a contiguous fp32 CUDA tensor in the torch weight layout")
    lib = _lib.lib()
    nbytes = lib.selavi_wgrad_workspace_bytes(geom.arr(0))
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _wgrad_ws[key] = ws
    with _Guard(x.device), _Prof("conv_wgrad", geom, "wgrad_kernel[tf32x3]"):
        _lib.check(lib.selavi_conv_wgrad(_lib.ptr(x), _lib.ptr(dz), _lib.ptr(dw), geom.arr(0), geom.ci, _lib.ptr(scale),
                                         _lib.ptr(shift), 1 if relu else 0, _lib.ptr(ws), 1 if accumulate else 0, passes,
                                         _lib.stream_ptr()), "selavi_conv_wgrad")
    return dw


# ---------------------------------------------------------------------------------------------- bf16x3 backward
def split_bf16(x, scale=None, shift=None, relu=False):
    """fp32 [.., Cs] -> (hi, lo) bf16 planes of act(x*scale+shift); hi = bf16(y), lo = bf16(y - hi)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("split_bf16 needs a contiguous fp32 CUDA tensor")
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with _Guard(x.device):
        _lib.check(_lib.lib().selavi_split_bf16(_lib.ptr(x), _lib.ptr(scale), _lib.ptr(shift), 1 if relu else 0, _lib.ptr(hi),
                                                _lib.ptr(lo), x.numel() // x.shape[-1], x.shape[-1], _lib.stream_ptr()),
                   "selavi_split_bf16")
    return hi, lo


def pack_weights_dgrad_bf16(w, geom, out=None):
    """torch weight [co, ci, taps...] -> bf16 hi/lo B operand of the data gradient (n = ci, k = (tap, co))."""
    w = w.detach()
    if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
        raise ValueError("weight must be a contiguous fp32 CUDA tensor")
    lib = _lib.lib()
    nbytes = lib.selavi_dgrad_wpack_bytes(geom.arr(1))
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    elif out.numel() != nbytes:
        raise ValueError("packed weight buffer has the wrong size")
    with _Guard(w.device):
        _lib.check(lib.selavi_dgrad_pack_weights(_lib.ptr(w), geom.arr(1), geom.co, _lib.ptr(out), _lib.stream_ptr()),
                   "selavi_dgrad_pack_weights")
    return out


def _chk_bf16(t, shape, name):
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
        raise ValueError(f"{name}: expected contiguous bf16 CUDA tensor of shape {tuple(shape)}")


def conv_dgrad_bf16(z_hi, z_lo, wpack_bf16, geom, out=None, accumulate=False, passes=3):
    _chk_bf16(z_hi, geom.out_shape(), "z_hi")
    _chk_bf16(z_lo, geom.out_shape(), "z_lo")
    if out is None:
        out = torch.empty(geom.in_shape(), dtype=torch.float32, device=z_hi.device)
        accumulate = False
    _chk(out, geom.in_shape(), "dx")
    with _Guard(z_hi.device), _Prof("conv_dgrad", geom, "dgrad_bf16_kernel"):
        _lib.check(_lib.lib().selavi_conv_dgrad_bf16(_lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(out), _lib.ptr(wpack_bf16),
                                                     geom.arr(1), 1 if accumulate else 0, passes, _lib.stream_ptr()),
                   "selavi_conv_dgrad_bf16")
    return out


def conv_wgrad_bf16_planes(a_hi, a_lo, z_hi, z_lo, geom, dw, accumulate=False, passes=3):
    """weight gradient with BOTH operands pre-split: a_hi/a_lo = bf16 planes of the conv input activation (emitted by the
    BatchNorm-backward apply pass of the unit that produced it), z_hi/z_lo = planes of the gradient wrt the conv output"""
    _chk_bf16(a_hi, geom.in_shape(), "a_hi")
    _chk_bf16(a_lo, geom.in_shape(), "a_lo")
    _chk_bf16(z_hi, geom.out_shape(), "z_hi")
    _chk_bf16(z_lo, geom.out_shape(), "z_lo")
    if not (dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == geom.co * geom.ci * geom.taps):
        raise ValueError("dw must be a contiguous fp32 CUDA tensor in the torch weight layout")
    lib = _lib.lib()
    nbytes = lib.selavi_wgrad_workspace_bytes(geom.arr(0))
    key = (a_hi.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a_hi.device)
        _wgrad_ws[key] = ws
    with _Guard(a_hi.device), _Prof("conv_wgrad", geom, "wgrad_bf16_kernel(+reduce)"):
        _lib.check(lib.selavi_conv_wgrad_bf16_planes(_lib.ptr(a_hi), _lib.ptr(a_lo), _lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(dw),
                                                     geom.arr(0), geom.ci, _lib.ptr(ws), 1 if accumulate else 0, passes,
                                                     _lib.stream_ptr()), "selavi_conv_wgrad_bf16_planes")
    return dw


def conv_wgrad_bf16(x, z_hi, z_lo, geom, dw, scale=None, shift=None, relu=False, accumulate=False, passes=3):
    _chk(x, geom.in_shape(), "x")
    _chk_bf16(z_hi, geom.out_shape(), "z_hi")
    _chk_bf16(z_lo, geom.out_shape(), "z_lo")
    if not (dw.is_cuda and dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == geom.co * geom.ci * geom.taps):
        raise ValueError("dw must be a contiguous fp32 CUDA tensor in the torch weight layout")
    lib = _lib.lib()
    nbytes = lib.selavi_wgrad_workspace_bytes(geom.arr(0))
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _wgrad_ws[key] = ws
    with _Guard(x.device), _Prof("conv_wgrad", geom, "wgrad_bf16_kernel(+split,reduce)"):
        _lib.check(lib.selavi_conv_wgrad_bf16(_lib.ptr(x), _lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(dw), geom.arr(0), geom.ci,
                                              _lib.ptr(scale), _lib.ptr(shift), 1 if relu else 0, _lib.ptr(ws),
                                              1 if accumulate else 0, passes, _lib.stream_ptr()), "selavi_conv_wgrad_bf16")
    return dw
