"""Two launches each of the bf16x3 weight-gradient kernel on the layer-1 shapes (for `ncu --set full -k regex:wgrad_bf16`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops

dev = torch.device("cuda:0")
for ci, co, k in ((64, 144, (1, 3, 3)), (144, 64, (3, 1, 1))):
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    geom = ops.ConvGeom(16, ci, co, (32, 56, 56), k, (1, 1, 1), p)
    x = torch.randn(geom.in_shape(), device=dev)
    z_hi, z_lo = ops.split_bf16(torch.randn(geom.out_shape(), device=dev))
    dw = torch.empty(co, ci, *k, device=dev)
    for _ in range(2):
        ops.conv_wgrad_bf16(x, z_hi, z_lo, geom, dw)
    torch.cuda.synchronize()
print("done")
