"""Tap-reuse data gradient with / without the fused BatchNorm-backward statistics (layer shapes of configs[1])."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops
from tools.quick_bench import timeit

dev = torch.device("cuda:0")
ops.FWD_KERNEL = "halo"
for name, nb, ci, co, thw, k in [("l1 spatial dz144->dx64", 16, 64, 144, (32, 56, 56), (1, 3, 3)),
                                 ("l1 temporal dz64->dx144", 16, 144, 64, (32, 56, 56), (3, 1, 1)),
                                 ("l2 spatial dz288->dx128", 16, 128, 288, (16, 28, 28), (1, 3, 3)),
                                 ("l2 temporal dz128->dx288", 16, 288, 128, (16, 28, 28), (3, 1, 1)),
                                 ("l3 spatial dz576->dx256", 16, 256, 576, (8, 14, 14), (1, 3, 3))]:
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    geom = ops.ConvGeom(nb, ci, co, thw, k, (1, 1, 1), p)
    w = torch.randn(co, ci, *k, device=dev) * 0.05
    dz = torch.randn(geom.out_shape(), device=dev)
    z_hi, z_lo = ops.split_bf16(dz)
    del dz
    wp = ops.pack_weights_halo(w, geom, mode=1)
    zprev = torch.randn(geom.in_shape(), device=dev)
    v = [torch.rand(geom.cis, device=dev) + 0.5 for _ in range(4)]
    dx = torch.empty(geom.in_shape(), device=dev)
    t0 = timeit(lambda: ops.conv_dgrad_halo(z_hi, z_lo, wp, geom, out=dx))
    t1 = timeit(lambda: ops.conv_dgrad_halo(z_hi, z_lo, wp, geom, bn=(zprev, *v)))
    # what the fusion replaces: the separate reduce pass over g and z
    from selavi_b200 import _lib
    lib = _lib.lib()
    M, cs = geom.m_in, geom.cis
    nblk = lib.selavi_bn_bwd_blocks(M)
    partial = torch.empty(nblk * 2 * cs, device=dev)
    sums = torch.empty(2 * cs, dtype=torch.float64, device=dev)
    t2 = timeit(lambda: _lib.check(lib.selavi_bn_bwd_reduce(_lib.ptr(dx), _lib.ptr(zprev), None, 2, _lib.ptr(v[0]), _lib.ptr(v[1]), _lib.ptr(v[2]),
                                                            _lib.ptr(v[3]), M, cs, _lib.ptr(partial), _lib.ptr(sums), _lib.stream_ptr()), "r"))
    print(f"{name}: dgrad {t0:.3f} ms | dgrad + fused stats {t1:.3f} ms (+{t1 - t0:.3f}) | separate bn_bwd_reduce {t2:.3f} ms", flush=True)
    del z_hi, z_lo, zprev, dx
