"""Where does the tap-reuse kernel's time go?  Times the layer-1 forward / data-gradient launches with parts of the
kernel switched off (SELAVI_HALO_FLAGS debug bits: 2 = weights loaded once, 4 = no activation gathers, 8 = no output
stores; 32 = CTA pairs (cta_group::2)).  Results of the debug variants are invalid by construction;
only their durations are read.  Not the bench contract."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
ops.FWD_KERNEL = "halo"


def timeit(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


VARIANTS = [(0, "single CTA"), (32, "CTA pair"), (2, "no B"), (4, "no A"), (8, "no store"), (14, "MMA only"), (32 | 14, "pair MMA only")]


def run(name, ci, co, thw, k, nb=16):
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    geom = ops.ConvGeom(nb, ci, co, thw, k, (1, 1, 1), p)
    flop = 2.0 * geom.m_out * co * ci * geom.taps
    w = torch.randn(co, ci, *k, device=dev) * 0.05
    x = torch.randn(geom.in_shape(), device=dev)
    sc, sf = torch.rand(geom.cis, device=dev) + 0.5, torch.randn(geom.cis, device=dev) * 0.3
    wp, st = ops.pack_weights_halo(w, geom), ops.stats_buffer(geom, dev, halo=True)
    y = torch.empty(geom.out_shape(), device=dev)
    z_hi, z_lo = ops.split_bf16(torch.randn(geom.out_shape(), device=dev))
    wpd = ops.pack_weights_halo(w, geom, mode=1)
    dx = torch.empty(geom.in_shape(), device=dev)
    ref = {}
    for kind in ("fwd", "dgrad"):
        row = []
        for flags, label in VARIANTS:
            ops.HALO_FLAGS = flags
            if kind == "fwd":
                ms = timeit(lambda: ops.conv_forward_halo(x, wp, geom, out=y, scale=sc, shift=sf, relu=True, stats=st))
                out = y
            else:
                ms = timeit(lambda: ops.conv_dgrad_halo(z_hi, z_lo, wpd, geom, out=dx))
                out = dx
            if flags == 0:
                ref[kind] = out.clone()
            if flags == 32:
                row.append(f"[pair vs single {float((out - ref[kind]).norm() / ref[kind].norm()):.1e}]")
            row.append(f"{label} {ms:.3f} ms ({flop / ms / 1e9:.0f} TF/s)")
        ops.HALO_FLAGS = 0
        print(f"{name} {kind}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    run("l1_spatial 64->144", 64, 144, (32, 56, 56), (1, 3, 3))
    run("l1_temporal 144->64", 144, 64, (32, 56, 56), (3, 1, 1))
    run("l2_spatial 128->288", 128, 288, (16, 28, 28), (1, 3, 3))
    run("l2_temporal 288->128", 288, 128, (16, 28, 28), (3, 1, 1))
