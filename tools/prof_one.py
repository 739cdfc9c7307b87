"""Launch one kernel shape a few times (for ncu): python tools/prof_one.py {wgrad|fwd|dgrad} [layer]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops
dev = torch.device("cuda:0")
what = sys.argv[1]
shapes = {"l1s": (16, 64, 144, (32, 56, 56), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          "l1t": (16, 144, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0))}
nb, ci, co, thw, k, s, p = shapes[sys.argv[2] if len(sys.argv) > 2 else "l1s"]
geom = ops.ConvGeom(nb, ci, co, thw, k, s, p)
x = torch.randn(geom.in_shape(), device=dev)
w = torch.randn(co, ci, *k, device=dev) * 0.05
dz = torch.randn(geom.out_shape(), device=dev)
for _ in range(3):
    if what == "wgrad":
        ops.conv_wgrad(x, dz, geom, torch.empty_like(w))
    elif what == "fwd":
        ops.conv_forward(x, ops.pack_weights(w, geom, 0), geom, stats=ops.stats_buffer(geom, dev))
    else:
        ops.conv_dgrad(dz, ops.pack_weights(w, geom, 1), geom)
torch.cuda.synchronize()
