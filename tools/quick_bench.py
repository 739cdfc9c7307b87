"""Quick device timings of individual kernels (CUDA events, warm-up, L2-exceeding inputs). Not the bench contract."""
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selavi_b200 import ops
from selavi_b200.sk_utils import SKWorkspace, sk_solve_raw

dev = torch.device("cuda:0")
torch.cuda.set_device(0)


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


def bench_sk(N=200000, K=309, iters=100):
    g = torch.Generator(device=dev).manual_seed(0)
    PS = torch.softmax(torch.randn(N, K, dtype=torch.float64, device=dev, generator=g), 1) * \
        torch.softmax(torch.randn(N, K, dtype=torch.float64, device=dev, generator=g), 1)
    ws = SKWorkspace(K, N, dev)
    sk_solve_raw(PS, N, 20.0, None, ws, max_iters=10, stop_on_converge=False, do_final=False)
    ms = timeit(lambda: sk_solve_raw(PS, N, 20.0, None, ws, max_iters=iters, stop_on_converge=False, do_prep=False, do_final=False))
    gbs = N * K * 8 * iters / (ms * 1e-3) / 1e9
    print(f"SK N={N} K={K}: {iters} iters in {ms:.3f} ms -> {iters / ms * 1e3:.0f} it/s, {gbs:.0f} GB/s algorithmic", flush=True)
    PS2 = PS.clone()
    t0 = time.time(); sk_solve_raw(PS2, N, 20.0, None, ws); torch.cuda.synchronize()
    print(f"SK full solve: iters={int(ws.iters.item())} err={float(ws.err.item()):.4f} wall={time.time() - t0:.4f}s", flush=True)


def bench_conv(name, nb, ci, co, thw, k, s, p):
    geom = ops.ConvGeom(nb, ci, co, thw, k, s, p)
    x = torch.randn(geom.in_shape(), device=dev)
    w = torch.randn(co, ci, *k, device=dev) * 0.05
    wp, wpt = ops.pack_weights(w, geom, 0), ops.pack_weights(w, geom, 1)
    y = torch.empty(geom.out_shape(), device=dev)
    dz = torch.randn(geom.out_shape(), device=dev)
    dx = torch.empty(geom.in_shape(), device=dev)
    dw = torch.empty_like(w)
    stats = ops.stats_buffer(geom, dev)
    flop = 2.0 * geom.m_out * co * ci * geom.taps
    for passes in (3, 1):
        ms = timeit(lambda: ops.conv_forward(x, wp, geom, out=y, stats=stats, passes=passes))
        msd = timeit(lambda: ops.conv_dgrad(dz, wpt, geom, out=dx, passes=passes))
        msw = timeit(lambda: ops.conv_wgrad(x, dz, geom, dw, passes=passes))
        z_hi, z_lo = ops.split_bf16(dz)
        wpb = ops.pack_weights_dgrad_bf16(w, geom)
        msdb = timeit(lambda: ops.conv_dgrad_bf16(z_hi, z_lo, wpb, geom, out=dx, passes=passes))
        mswb = timeit(lambda: ops.conv_wgrad_bf16(x, z_hi, z_lo, geom, dw, passes=passes))
        print(f"{name} passes={passes}: fwd {ms:.3f} ms ({flop / ms / 1e9:.1f} TF/s)  dgrad {msd:.3f} ms ({flop / msd / 1e9:.1f})"
              f"  wgrad {msw:.3f} ms ({flop / msw / 1e9:.1f}) | bf16 presplit: dgrad {msdb:.3f} ms ({flop / msdb / 1e9:.1f})"
              f"  wgrad {mswb:.3f} ms ({flop / mswb / 1e9:.1f})", flush=True)


def bench_halo(name, nb, ci, co, thw, k):
    """tap-reuse fp16x3 forward vs the tf32x3 implicit GEMM on the same layer (with the fused BN+ReLU prologue)"""
    ops.FWD_KERNEL = "halo"
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    geom = ops.ConvGeom(nb, ci, co, thw, k, (1, 1, 1), p)
    x = torch.randn(geom.in_shape(), device=dev)
    w = torch.randn(co, ci, *k, device=dev) * 0.05
    sc = torch.rand(geom.cis, device=dev) + 0.5
    sf = torch.randn(geom.cis, device=dev) * 0.3
    y = torch.empty(geom.out_shape(), device=dev)
    flop = 2.0 * geom.m_out * co * ci * geom.taps
    wp = ops.pack_weights(w, geom, 0)
    st = ops.stats_buffer(geom, dev)
    ms0 = timeit(lambda: ops.conv_forward(x, wp, geom, out=y, scale=sc, shift=sf, relu=True, stats=st, passes=3))
    y0 = y.clone()
    plan = ops.halo_plan(geom)
    wph = ops.pack_weights_halo(w, geom)
    sth = ops.stats_buffer(geom, dev, halo=True)
    ms1 = timeit(lambda: ops.conv_forward_halo(x, wph, geom, out=y, scale=sc, shift=sf, relu=True, stats=sth))
    rel = float((y - y0).norm() / y0.norm())
    # data gradient: cp.async-fed bf16x3 implicit GEMM vs the same through the tap-reuse kernel
    dz = torch.randn(geom.out_shape(), device=dev)
    z_hi, z_lo = ops.split_bf16(dz)
    dx = torch.empty(geom.in_shape(), device=dev)
    wpb = ops.pack_weights_dgrad_bf16(w, geom)
    msd0 = timeit(lambda: ops.conv_dgrad_bf16(z_hi, z_lo, wpb, geom, out=dx))
    dx0 = dx.clone()
    wpd = ops.pack_weights_halo(w, geom, mode=1)
    msd1 = timeit(lambda: ops.conv_dgrad_halo(z_hi, z_lo, wpd, geom, out=dx))
    reld = float((dx - dx0).norm() / dx0.norm())
    print(f"{name}: dgrad bf16x3 igemm {msd0:.3f} ms ({flop / msd0 / 1e9:.1f} TF/s) | halo {msd1:.3f} ms ({flop / msd1 / 1e9:.1f} TF/s) "
          f"rel diff {reld:.2e}", flush=True)
    byts = (x.numel() + y.numel()) * 4
    print(f"{name}: igemm tf32x3 {ms0:.3f} ms ({flop / ms0 / 1e9:.1f} TF/s) | halo fp16x3 {ms1:.3f} ms ({flop / ms1 / 1e9:.1f} TF/s, "
          f"{byts / ms1 / 1e6:.0f} GB/s in+out) plan(m_tiles,bnt,nt,wbytes)={plan} rel diff {rel:.2e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "sk"):
        bench_sk()
    if what == "halo":
        bench_halo("l1_temporal 144->64", 16, 144, 64, (32, 56, 56), (3, 1, 1))
        bench_halo("l1_spatial 64->144", 16, 64, 144, (32, 56, 56), (1, 3, 3))
        bench_halo("stem_t 45->64", 16, 45, 64, (32, 56, 56), (3, 1, 1))
        bench_halo("l2_spatial 128->230", 16, 128, 230, (16, 28, 28), (1, 3, 3))
        bench_halo("l2_temporal 230->128", 16, 230, 128, (16, 28, 28), (3, 1, 1))
        bench_halo("l2_spatial 128->288", 16, 128, 288, (16, 28, 28), (1, 3, 3))
        bench_halo("l2_temporal 288->128", 16, 288, 128, (16, 28, 28), (3, 1, 1))
        bench_halo("l3_spatial 256->576", 16, 256, 576, (8, 14, 14), (1, 3, 3))
        bench_halo("l3_temporal 576->256", 16, 576, 256, (8, 14, 14), (3, 1, 1))
        bench_halo("l4_spatial 512->1152", 16, 512, 1152, (4, 7, 7), (1, 3, 3))
        bench_halo("l4_temporal 1152->512", 16, 1152, 512, (4, 7, 7), (3, 1, 1))
        bench_halo("audio_l1 64->64", 16, 64, 64, (1, 65, 50), (1, 3, 3))
    if what in ("all", "conv"):
        bench_conv("l1_spatial", 16, 64, 144, (32, 56, 56), (1, 3, 3), (1, 1, 1), (0, 1, 1))
        bench_conv("l1_temporal", 16, 144, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0))
        bench_conv("l2_spatial", 16, 128, 288, (16, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1))
        bench_conv("l4_spatial", 16, 512, 1152, (4, 7, 7), (1, 3, 3), (1, 1, 1), (0, 1, 1))
