"""GPU parity of the tcgen05 implicit-GEMM convolution kernels (forward, data gradient, weight gradient)
against torch's float64 convolution on the same inputs (a floating-point kernel: torch reference, tolerance
stated per test).  Shapes are the layer shapes of R(2+1)D-18 / ResNet-9, reduced batch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# name: (nb, ci, co, (T,H,W), kernel, stride, padding)
LAYERS = {
    "gemm_1x1x1": (2, 64, 64, (2, 8, 8), (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    "v_stem0_7x7": (1, 3, 45, (4, 32, 32), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    "v_stem3_t": (1, 45, 64, (4, 16, 16), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    "v_l1_spatial": (1, 64, 144, (4, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    "v_l1_temporal": (1, 144, 64, (4, 14, 14), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    "v_l2_spatial_s2": (1, 64, 230, (4, 28, 28), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    "v_l2_temporal_s2": (1, 230, 128, (8, 14, 14), (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    "v_l2_downsample": (1, 64, 128, (8, 28, 28), (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    "v_l3_spatial_288": (1, 128, 288, (2, 14, 14), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    "v_l4_spatial_921": (2, 256, 921, (2, 14, 14), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    "v_l4_temporal_1152": (2, 1152, 512, (4, 7, 7), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    "a_conv1_7x7": (2, 1, 64, (1, 65, 50), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    "a_l2_3x3_s2": (2, 64, 128, (1, 33, 25), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    "ragged_m": (1, 12, 20, (3, 5, 7), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
}


def _mk(name, device):
    from selavi_b200 import ops
    nb, ci, co, thw, k, s, p = LAYERS[name]
    g = torch.Generator(device=device).manual_seed(hash(name) % 1000)
    x = torch.randn(nb, ci, *thw, device=device, generator=g)
    w = torch.randn(co, ci, *k, device=device, generator=g) * (1.0 / (ci * k[0] * k[1] * k[2]) ** 0.5)
    geom = ops.ConvGeom(nb, ci, co, thw, k, s, p)
    return x, w, geom, (s, p)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("name", sorted(LAYERS))
# passes=3 (tf32x3): operand error ~2^-21; the tensor core adds each K=8 partial product into the fp32
# accumulator with truncation, so the error grows ~n_mma * 2^-25 (measured 8e-7 at K=64, 4e-6 at K=576).
# passes=6 (fp16x3, the default forward of the strided / 7x7 / 1x1 convolutions): fp16 hi/lo operands, same accuracy class.
@pytest.mark.parametrize("passes,tol", [(6, 5e-5), (3, 5e-5), (1, 2e-3)])
def test_conv_forward(cuda_device, name, passes, tol):
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    ref = F.conv3d(x.double(), w.double(), None, s, p)
    y = ops.conv_forward(ops.to_channels_last(x), ops.pack_weights(w, geom, 2 if passes == 6 else 0), geom, passes=passes)
    assert y[..., geom.co:].abs().max().item() == 0 if geom.cos > geom.co else True
    err = _rel(ops.from_channels_last(y, geom.co), ref)
    print(f"{name} fwd passes={passes} rel={err:.3e}")
    assert err < tol


@pytest.mark.parametrize("passes", [6, 3])
@pytest.mark.parametrize("name", ["v_l1_spatial", "v_stem3_t", "a_l2_3x3_s2", "ragged_m", "v_l2_temporal_s2", "v_stem0_7x7"])
def test_conv_forward_fused_bn_relu_prologue_and_stats(cuda_device, name, passes):
    """prologue = train-mode BN(+ReLU) of the previous layer applied on the fly; epilogue = per-channel stats."""
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    scale = torch.rand(geom.cis, device=cuda_device, generator=g) + 0.5
    shift = torch.randn(geom.cis, device=cuda_device, generator=g) * 0.3
    xn = torch.relu(x.double() * scale[:geom.ci].double().view(1, -1, 1, 1, 1) + shift[:geom.ci].double().view(1, -1, 1, 1, 1))
    ref = F.conv3d(xn, w.double(), None, s, p)
    stats = ops.stats_buffer(geom, cuda_device)
    y = ops.conv_forward(ops.to_channels_last(x), ops.pack_weights(w, geom, 2 if passes == 6 else 0), geom, scale=scale, shift=shift,
                         relu=True, stats=stats, passes=passes)
    assert _rel(ops.from_channels_last(y, geom.co), ref) < 5e-5
    tot = stats.double().sum(0)[:, :geom.co]
    # fp32 per-tile partial sums: tolerance relative to the L1 / L2 mass of the channel
    l1 = ref.abs().sum((0, 2, 3, 4))
    assert float(((tot[0] - ref.sum((0, 2, 3, 4))).abs() / l1).max()) < 1e-5
    assert float(((tot[1] - (ref * ref).sum((0, 2, 3, 4))).abs() / (ref * ref).sum((0, 2, 3, 4))).max()) < 1e-4


@pytest.mark.parametrize("name", sorted(LAYERS))
def test_conv_dgrad(cuda_device, name):
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    xd = x.double().requires_grad_(True)
    ref_y = F.conv3d(xd, w.double(), None, s, p)
    dz = torch.randn(ref_y.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(7))
    ref_dx, = torch.autograd.grad(ref_y, xd, dz.double())
    dx = ops.conv_dgrad(ops.to_channels_last(dz), ops.pack_weights(w, geom, 1), geom)
    err = _rel(ops.from_channels_last(dx, geom.ci), ref_dx)
    print(f"{name} dgrad rel={err:.3e}")
    assert err < 5e-5
    # accumulate=True adds onto the existing gradient (residual joins)
    dx2 = ops.conv_dgrad(ops.to_channels_last(dz), ops.pack_weights(w, geom, 1), geom, out=dx.clone(), accumulate=True)
    assert _rel(dx2, 2 * dx) < 1e-6


# passes 3 = bf16x3 MN-major kernel (default), 13 = tf32x3 K-major transposing kernel
@pytest.mark.parametrize("passes", [3, 13])
@pytest.mark.parametrize("name", sorted(LAYERS))
def test_conv_wgrad(cuda_device, name, passes):
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    wd = w.double().requires_grad_(True)
    ref_y = F.conv3d(x.double(), wd, None, s, p)
    dz = torch.randn(ref_y.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(9))
    ref_dw, = torch.autograd.grad(ref_y, wd, dz.double())
    dw = torch.empty_like(w)
    ops.conv_wgrad(ops.to_channels_last(x), ops.to_channels_last(dz), geom, dw, passes=passes)
    err = _rel(dw, ref_dw)
    print(f"{name} wgrad passes={passes} rel={err:.3e}")
    assert err < 5e-5


@pytest.mark.parametrize("name", sorted(LAYERS))
def test_conv_dgrad_bf16x3(cuda_device, name):
    """data gradient on pre-split bf16 hi/lo planes (cp.async-fed kind::f16 kernel), tolerance 5e-5 like tf32x3"""
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    xd = x.double().requires_grad_(True)
    ref_y = F.conv3d(xd, w.double(), None, s, p)
    dz = torch.randn(ref_y.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(7))
    ref_dx, = torch.autograd.grad(ref_y, xd, dz.double())
    z_hi, z_lo = ops.split_bf16(ops.to_channels_last(dz))
    torch.testing.assert_close(z_hi.float() + z_lo.float(), ops.to_channels_last(dz), rtol=2e-5, atol=1e-30)
    wp = ops.pack_weights_dgrad_bf16(w, geom)
    dx = ops.conv_dgrad_bf16(z_hi, z_lo, wp, geom)
    err = _rel(ops.from_channels_last(dx, geom.ci), ref_dx)
    print(f"{name} dgrad bf16x3 rel={err:.3e}")
    assert err < 5e-5
    dx2 = ops.conv_dgrad_bf16(z_hi, z_lo, wp, geom, out=dx.clone(), accumulate=True)
    assert _rel(dx2, 2 * dx) < 1e-6
    # wgrad on the same pre-split planes
    wd = w.double().requires_grad_(True)
    ref_dw, = torch.autograd.grad(F.conv3d(x.double(), wd, None, s, p), wd, dz.double())
    dw = torch.empty_like(w)
    ops.conv_wgrad_bf16(ops.to_channels_last(x), z_hi, z_lo, geom, dw)
    assert _rel(dw, ref_dw) < 5e-5


# ------------------------------------------------------------------------------------------------ tap-reuse forward
# name: (nb, ci, co, (T,H,W), kernel): stride 1, padding 1 along the 3-wide dims (the halo kernel's domain)
HALO_LAYERS = {
    "t_l1": (1, 144, 64, (8, 28, 28), (3, 1, 1)),            # 2.25 K chunks, N=64
    "t_ragged": (2, 45, 64, (5, 9, 7), (3, 1, 1)),           # T < 8, S % 16 != 0, one partial K chunk
    "t_l2_230": (1, 230, 128, (16, 14, 14), (3, 1, 1)),      # cs=232: last k16 half valid, two frame blocks
    "t_l3_two_ntiles": (1, 64, 460, (8, 8, 8), (3, 1, 1)),   # two N tiles
    "s_l1": (1, 64, 144, (2, 56, 56), (1, 3, 3)),            # WP=58: 246 slot rows
    "s_ragged": (2, 24, 40, (3, 11, 13), (1, 3, 3)),         # last tile of the frame partial, cs=24
    "s_l2_128_230": (1, 128, 230, (2, 28, 28), (1, 3, 3)),   # two K chunks, cd=232
    "s_audio_2d": (2, 64, 64, (1, 65, 50), (1, 3, 3)),       # 2-D conv (T=1)
}


def _mk_halo(name, device):
    from selavi_b200 import ops
    nb, ci, co, thw, k = HALO_LAYERS[name]
    p = (1, 0, 0) if k[0] == 3 else (0, 1, 1)
    g = torch.Generator(device=device).manual_seed(hash(name) % 1000)
    x = torch.randn(nb, ci, *thw, device=device, generator=g)
    w = torch.randn(co, ci, *k, device=device, generator=g) * (1.0 / (ci * k[0] * k[1] * k[2]) ** 0.5)
    return x, w, ops.ConvGeom(nb, ci, co, thw, k, (1, 1, 1), p), p


@pytest.mark.parametrize("name", sorted(HALO_LAYERS))
def test_conv_forward_halo(cuda_device, name, monkeypatch):
    """fp16x3 tap-reuse kernel vs float64 conv: same 5e-5 bar as the tf32x3 kernel (22-bit operands, fp32 accumulate)."""
    from selavi_b200 import ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    x, w, geom, p = _mk_halo(name, cuda_device)
    assert ops.halo_plan(geom) is not None
    ref = F.conv3d(x.double(), w.double(), None, 1, p)
    y = ops.conv_forward_halo(ops.to_channels_last(x), ops.pack_weights_halo(w, geom), geom)
    if geom.cos > geom.co:
        assert y[..., geom.co:].abs().max().item() == 0
    err = _rel(ops.from_channels_last(y, geom.co), ref)
    print(f"{name} halo fwd rel={err:.3e}")
    assert err < 5e-5


@pytest.mark.parametrize("name", sorted(HALO_LAYERS))
def test_conv_forward_halo_prologue_and_stats(cuda_device, name, monkeypatch):
    from selavi_b200 import ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    x, w, geom, p = _mk_halo(name, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    scale = torch.rand(geom.cis, device=cuda_device, generator=g) + 0.5
    shift = torch.randn(geom.cis, device=cuda_device, generator=g) * 0.3
    xn = torch.relu(x.double() * scale[:geom.ci].double().view(1, -1, 1, 1, 1) + shift[:geom.ci].double().view(1, -1, 1, 1, 1))
    ref = F.conv3d(xn, w.double(), None, 1, p)
    stats = ops.stats_buffer(geom, cuda_device, halo=True)
    y = ops.conv_forward_halo(ops.to_channels_last(x), ops.pack_weights_halo(w, geom), geom, scale=scale, shift=shift,
                              relu=True, stats=stats)
    assert _rel(ops.from_channels_last(y, geom.co), ref) < 5e-5
    tot = stats.double().sum(0)[:, :geom.co]
    l1 = ref.abs().sum((0, 2, 3, 4))
    assert float(((tot[0] - ref.sum((0, 2, 3, 4))).abs() / l1).max()) < 1e-5
    assert float(((tot[1] - (ref * ref).sum((0, 2, 3, 4))).abs() / (ref * ref).sum((0, 2, 3, 4))).max()) < 1e-4


def test_halo_plan_rejects_other_geometries(cuda_device, monkeypatch):
    from selavi_b200 import ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    for name in ("v_stem0_7x7", "v_l2_spatial_s2", "v_l2_temporal_s2", "v_l2_downsample", "ragged_m"):
        nb, ci, co, thw, k, s, p = LAYERS[name]
        assert ops.halo_plan(ops.ConvGeom(nb, ci, co, thw, k, s, p)) is None


@pytest.mark.parametrize("name", sorted(HALO_LAYERS))
def test_conv_dgrad_halo(cuda_device, name, monkeypatch):
    """data gradient through the tap-reuse kernel (bf16x3 on pre-split planes, flipped taps) vs float64 autograd"""
    from selavi_b200 import ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    x, w, geom, p = _mk_halo(name, cuda_device)
    assert ops.halo_plan(geom, 1) is not None
    xd = x.double().requires_grad_(True)
    ref_y = F.conv3d(xd, w.double(), None, 1, p)
    dz = torch.randn(ref_y.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(7))
    ref_dx, = torch.autograd.grad(ref_y, xd, dz.double())
    z_hi, z_lo = ops.split_bf16(ops.to_channels_last(dz))
    wp = ops.pack_weights_halo(w, geom, mode=1)
    dx = ops.conv_dgrad_halo(z_hi, z_lo, wp, geom)
    if geom.cis > geom.ci:
        assert dx[..., geom.ci:].abs().max().item() == 0
    err = _rel(ops.from_channels_last(dx, geom.ci), ref_dx)
    print(f"{name} halo dgrad rel={err:.3e}")
    assert err < 5e-5
    dx2 = ops.conv_dgrad_halo(z_hi, z_lo, wp, geom, out=dx.clone(), accumulate=True)
    assert _rel(dx2, 2 * dx) < 1e-6


@pytest.mark.parametrize("name", sorted(HALO_LAYERS))
def test_conv_dgrad_halo_fused_bn_backward_stats(cuda_device, name, monkeypatch):
    """the data-gradient epilogue also emits sum(g*m), sum(g*m*zhat) of the BatchNorm-backward pass that consumes g
    (m = relu mask of the producing unit, zhat its normalised output): same dx bit for bit, sums vs float64 torch, and
    the activation planes emitted by bn_bwd_apply equal split_bf16 of the activation"""
    from selavi_b200 import _lib, ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    x, w, geom, p = _mk_halo(name, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(11)
    dz = torch.randn(geom.out_shape(), device=cuda_device, generator=g)
    dz[..., geom.co:] = 0
    z_hi, z_lo = ops.split_bf16(dz)
    wp = ops.pack_weights_halo(w, geom, mode=1)
    dx_ref = ops.conv_dgrad_halo(z_hi, z_lo, wp, geom)
    cs = geom.cis
    zprev = torch.randn(geom.in_shape(), device=cuda_device, generator=g)
    zprev[..., geom.ci:] = 0
    scale = torch.rand(cs, device=cuda_device, generator=g) + 0.5
    shift = torch.randn(cs, device=cuda_device, generator=g) * 0.3
    mean = torch.randn(cs, device=cuda_device, generator=g) * 0.2
    invstd = torch.rand(cs, device=cuda_device, generator=g) + 0.5
    for v in (scale, shift, mean, invstd):
        v[geom.ci:] = 0
    dx, stats = ops.conv_dgrad_halo(z_hi, z_lo, wp, geom, bn=(zprev, scale, shift, mean, invstd))
    assert torch.equal(dx, dx_ref)
    tot = stats.double().sum(0)[:, :cs]
    m = (zprev.double() * scale.double() + shift.double()) > 0
    gm = dx_ref.double() * m
    ref1 = gm.reshape(-1, cs).sum(0)
    ref2 = (gm * (zprev.double() - mean.double()) * invstd.double()).reshape(-1, cs).sum(0)
    l1 = gm.abs().reshape(-1, cs).sum(0) + 1e-30
    l2 = (gm * (zprev.double() - mean.double()) * invstd.double()).abs().reshape(-1, cs).sum(0) + 1e-30
    assert float(((tot[0] - ref1).abs() / l1).max()) < 1e-5
    assert float(((tot[1] - ref2).abs() / l2).max()) < 1e-5
    # activation planes from the BatchNorm-backward apply pass == split_bf16 of relu(z*scale+shift)
    sums = torch.zeros(2 * cs, dtype=torch.float64, device=cuda_device)
    M = zprev.numel() // cs
    d_hi, d_lo, a_hi, a_lo = (torch.empty(zprev.shape, dtype=torch.bfloat16, device=cuda_device) for _ in range(4))
    _lib.check(_lib.lib().selavi_bn_bwd_apply(_lib.ptr(dx), _lib.ptr(zprev), None, 2, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(mean),
                                              _lib.ptr(invstd), _lib.ptr(sums), float(M), M, cs, None, None, 0, _lib.ptr(d_hi),
                                              _lib.ptr(d_lo), _lib.ptr(a_hi), _lib.ptr(a_lo), _lib.stream_ptr()), "bn_bwd_apply")
    e_hi, e_lo = ops.split_bf16(zprev, scale=scale, shift=shift, relu=True)
    assert torch.equal(a_hi, e_hi) and torch.equal(a_lo, e_lo)


@pytest.mark.parametrize("name", sorted(HALO_LAYERS))
def test_conv_halo_cta_pair_matches_single_cta(cuda_device, name, monkeypatch):
    """CTA-pair variant (cluster of 2, tcgen05 cta_group::2, each CTA stages half of every weight tile): the MMAs see
    the same operands in the same order, so forward (with prologue + stats) and data gradient are bit-identical to the
    single-CTA kernel, including an odd number of pixel tiles (the pair's second CTA recomputes and discards)."""
    from selavi_b200 import ops
    monkeypatch.setattr(ops, "FWD_KERNEL", "halo")
    x, w, geom, p = _mk_halo(name, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    scale = torch.rand(geom.cis, device=cuda_device, generator=g) + 0.5
    shift = torch.randn(geom.cis, device=cuda_device, generator=g) * 0.3
    x_cl, wp = ops.to_channels_last(x), ops.pack_weights_halo(w, geom)
    z_hi, z_lo = ops.split_bf16(torch.randn(geom.out_shape(), device=cuda_device, generator=g))
    wpd = ops.pack_weights_halo(w, geom, mode=1)
    out = {}
    for flags in (0, 32):
        monkeypatch.setattr(ops, "HALO_FLAGS", flags)
        stats = ops.stats_buffer(geom, cuda_device, halo=True).zero_()
        y = ops.conv_forward_halo(x_cl, wp, geom, scale=scale, shift=shift, relu=True, stats=stats)
        dx = ops.conv_dgrad_halo(z_hi, z_lo, wpd, geom)
        torch.cuda.synchronize()
        out[flags] = (y, stats, dx)
    for a, b in zip(out[0], out[32]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["v_l1_temporal", "v_stem3_t", "v_l4_temporal_1152", "v_l3_spatial_288", "ragged_m"])
def test_conv_wgrad_exchanged_operands(cuda_device, name, monkeypatch):
    """bf16x3 weight gradient with the operands exchanged (dz as the tap-shifted row operand, picked by the cost model
    for some stride-1 convs) against the plain orientation and the float64 reference."""
    from selavi_b200 import ops
    x, w, geom, (s, p) = _mk(name, cuda_device)
    wd = w.double().requires_grad_(True)
    ref_y = F.conv3d(x.double(), wd, None, s, p)
    dz = torch.randn(ref_y.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(11))
    ref_dw, = torch.autograd.grad(ref_y, wd, dz.double())
    z_hi, z_lo = ops.split_bf16(ops.to_channels_last(dz))
    x_cl = ops.to_channels_last(x)
    dws = []
    for noswap in ("", "1"):
        if noswap:
            monkeypatch.setenv("SELAVI_WGRAD_NOSWAP", "1")
        else:
            monkeypatch.delenv("SELAVI_WGRAD_NOSWAP", raising=False)
        dw = torch.full_like(w, float("nan"))
        ops.conv_wgrad_bf16(x_cl, z_hi, z_lo, geom, dw)
        assert _rel(dw, ref_dw) < 5e-5
        dws.append(dw)
    assert _rel(dws[0], dws[1]) < 2e-5
