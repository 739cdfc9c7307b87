"""Device timings of the HBM-bound elementwise kernels at layer-1 size (for roofline checks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.lib()


def timeit(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for cs in (144, 64, 232):
    M = 16 * 32 * 56 * 56 if cs != 232 else 16 * 32 * 28 * 28
    g = torch.randn(M, cs, device=dev); z = torch.randn(M, cs, device=dev)
    scale = torch.rand(cs, device=dev) + 0.5; shift = torch.randn(cs, device=dev); mean = torch.randn(cs, device=dev); invstd = torch.rand(cs, device=dev) + 0.5
    nblk = lib.selavi_bn_bwd_blocks(M)
    partial = torch.empty(nblk * 2 * cs, device=dev); sums = torch.empty(2 * cs, dtype=torch.float64, device=dev)
    zhi = torch.empty(M, cs, dtype=torch.bfloat16, device=dev); zlo = torch.empty_like(zhi)
    P = _lib.ptr
    red = lambda: lib.selavi_bn_bwd_reduce(P(g), P(z), None, 2, P(scale), P(shift), P(mean), P(invstd), M, cs, P(partial), P(sums), _lib.stream_ptr())
    app = lambda: lib.selavi_bn_bwd_apply(P(g), P(z), None, 2, P(scale), P(shift), P(mean), P(invstd), P(sums), float(M), M, cs, None, None, 0, P(zhi), P(zlo), _lib.stream_ptr())
    spl = lambda: ops.split_bf16(z, scale, shift, True)
    out = torch.empty_like(z)
    bna = lambda: lib.selavi_bn_apply(P(z), P(scale), P(shift), P(g), None, None, 1, P(out), M, cs, _lib.stream_ptr())
    gb = M * cs * 4 / 1e9
    for name, fn, nbytes in (("bn_bwd_reduce", red, 2 * gb), ("bn_bwd_apply", app, 3 * gb), ("split_bf16", spl, 2 * gb), ("bn_apply+res", bna, 3 * gb)):
        ms = timeit(fn)
        print(f"cs={cs:4d} M={M}: {name:14s} {ms:7.3f} ms  {nbytes / ms * 1e3:7.0f} GB/s", flush=True)
