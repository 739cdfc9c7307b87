"""One launch of the tap-reuse kernel per configuration, for `ncu --set full -k regex:conv_halo` captures:
launch order = fwd l1_spatial, fwd l1_temporal, dgrad l1_spatial, dgrad l1_temporal (each preceded by one warm-up launch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops

dev = torch.device("cuda:0")
ops.FWD_KERNEL = "halo"
CASES = [("fwd", 64, 144, (1, 3, 3)), ("fwd", 144, 64, (3, 1, 1)), ("dgrad", 64, 144, (1, 3, 3)), ("dgrad", 144, 64, (3, 1, 1))]
for kind, ci, co, k in CASES:
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    geom = ops.ConvGeom(16, ci, co, (32, 56, 56), k, (1, 1, 1), p)
    w = torch.randn(co, ci, *k, device=dev) * 0.05
    if kind == "fwd":
        x = torch.randn(geom.in_shape(), device=dev)
        sc, sf = torch.rand(geom.cis, device=dev) + 0.5, torch.randn(geom.cis, device=dev) * 0.3
        wp, st = ops.pack_weights_halo(w, geom), ops.stats_buffer(geom, dev, halo=True)
        y = torch.empty(geom.out_shape(), device=dev)
        for _ in range(2):
            ops.conv_forward_halo(x, wp, geom, out=y, scale=sc, shift=sf, relu=True, stats=st)
    else:
        z_hi, z_lo = ops.split_bf16(torch.randn(geom.out_shape(), device=dev))
        wp = ops.pack_weights_halo(w, geom, mode=1)
        dx = torch.empty(geom.in_shape(), device=dev)
        for _ in range(2):
            ops.conv_dgrad_halo(z_hi, z_lo, wp, geom, out=dx)
    torch.cuda.synchronize()
print("done")
