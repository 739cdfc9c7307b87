"""Diagnostics library (tcgen05 operand-layout probe), built apart from the product C ABI.

    from tools.probe import probe_lib;  lib = probe_lib()      # builds tools/probe/libselavi_probe.so on first use
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libselavi_probe.so")
_h = None


def probe_lib():
    global _h
    if _h is None:
        src = os.path.join(HERE, "probe.cu")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
                                   "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", src, "-o", LIB, "-lcudart"])
        _h = ctypes.CDLL(LIB)
        c_void_p, c_int, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong
        _h.selavi_debug_umma_probe.restype = c_int
        _h.selavi_debug_umma_probe.argtypes = [c_void_p, c_int, c_void_p, c_int, u64, u64, ctypes.c_uint, c_int, c_void_p, c_void_p,
                                               c_int, c_int, c_void_p, c_void_p]
    return _h
