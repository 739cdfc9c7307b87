"""GPU parity of the full model path (towers + heads + loss, forward and backward) against golden vectors produced
by the UNMODIFIED reference on CPU (tests/golden/model_*.npz, generator tests/golden/gen_golden_model.py).
Tolerance: BASELINE.json north_star — logits within 1e-3 relative of the reference fp32 logits."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_golden_model import CONFIGS, build, make_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_train_step_matches_reference(cuda_device, name):
    from selavi_b200 import model as sv_model
    from selavi_b200.utils import get_loss
    B, T, HW, ST, K, hc = CONFIGS[name]
    gold = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    video, spec, labels = make_inputs(name)
    m = build(sv_model.load_model, name).to(cuda_device).train()
    fv, fa = m(torch.from_numpy(video).to(cuda_device), torch.from_numpy(spec).to(cuda_device))
    lab = torch.from_numpy(labels).to(cuda_device)
    lab = lab[:, 0] if hc == 1 else lab
    loss = 0.5 * get_loss(fv, lab, headcount=hc) + 0.5 * get_loss(fa, lab, headcount=hc)
    m.zero_grad()
    loss.backward()
    lv = (torch.stack(list(fv)) if hc > 1 else fv[None]).detach().cpu().numpy()
    la = (torch.stack(list(fa)) if hc > 1 else fa[None]).detach().cpu().numpy()
    ev, ea = _rel(lv, gold["logits_v"]), _rel(la, gold["logits_a"])
    print(f"{name}: logits rel err video {ev:.2e} audio {ea:.2e} (vs fp64: {_rel(lv, gold['logits_v64']):.2e} "
          f"{_rel(la, gold['logits_a64']):.2e}; reference fp32 vs fp64: {_rel(gold['logits_v'], gold['logits_v64']):.2e})")
    assert ev < 1e-3 and ea < 1e-3
    assert np.array_equal(lv.argmax(-1), gold["logits_v"].argmax(-1))        # bit-exact argmax cluster ids
    assert np.array_equal(la.argmax(-1), gold["logits_a"].argmax(-1))
    assert abs(float(loss) - float(gold["loss"])) < 1e-4 * abs(float(gold["loss"]))
    # gradients: every parameter's norm, and full tensors for a sample of small parameters
    # cfg1 (B=2) puts BatchNorm1d on a batch of two near-identical clips: d(output)/d(input) ~ 1/|z0-z1| amplifies
    # rounding ~100x in both directions (reference fp32 vs fp64 already differ by 6e-4 there); mini_cfg2 is well conditioned
    gtol = 2e-2 if name == "cfg1" else 1e-2
    # Bar per parameter: max(gtol, 10 x the reference's OWN fp32 error against float64).  Gradients of BatchNorm biases are
    # sums over up to 25 M elements with heavy cancellation after up to 37 BatchNorm backward passes; the reference's fp32
    # result is itself 1e-3 .. 1e-2 off float64 there.  The forward operands here carry 22 significant bits against fp32's
    # 24 (logits 6e-5 vs the reference's 9e-6 at the benchmarked shape) and that ratio shows in the gradients: measured at
    # big_cfg2 (profiles/r02_precision_modes.txt) 2 of 247 norms above 5e-3, the worst 8.1e-3 = 7.7x the reference's own
    # 1.06e-3, identical for the bf16x3 and the tf32x3 backward (i.e. it is the saved forward activations, not the
    # backward operand width); at the batch-16 x 8-frame shapes the BatchNorm scale/bias norms sit at 3e-3 .. 7e-3.
    # gtol is therefore 1e-2 (every parameter's gradient norm within 1 % of float64), 2e-2 for the ill-conditioned cfg1.
    grads = {n: p.grad for n, p in m.named_parameters()}
    rows = []
    for n, ref_norm, ref_norm64 in zip(gold["grad_names"], gold["grad_norms"], gold["grad_norms64"]):
        g = grads[str(n)]
        assert g is not None, n
        e = abs(float(g.norm()) - ref_norm64) / max(ref_norm64, 1e-12)
        e32 = abs(ref_norm - ref_norm64) / max(ref_norm64, 1e-12)
        rows.append((e / max(gtol, 10 * e32), e, e32, str(n)))
    rows.sort(reverse=True)
    print(f"{name}: gradient norms vs float64, worst relative to their bar:")
    for r, e, e32, n in rows[:6]:
        print(f"    {n}: ours {e:.2e}, reference fp32 {e32:.2e}, bar {max(gtol, 10 * e32):.2e}")
    worst = max(e for _, e, _, _ in rows)
    assert rows[0][0] < 1.0, rows[0]
    trows = []
    for key in gold.files:
        if key.startswith("grad64/"):
            n = key[len("grad64/"):]
            e = _rel(grads[n].detach().cpu().numpy(), gold[key])
            # some gradients (first conv after 37 BN layers) are ill-conditioned: the reference's own fp32 result is
            # 6e-3 off its fp64 result there; allow 8x the reference's own error (tf32x3 carries ~2^-21 per operand)
            ref_err = _rel(gold["grad/" + n], gold[key])
            trows.append((e / max(2.5e-2 if name != 'cfg1' else 1e-1, 8 * ref_err), e, ref_err, n))
    trows.sort(reverse=True)
    print(f"{name}: full gradient tensors vs float64, worst relative to their bar:")
    for r, e, e32, n in trows[:6]:
        print(f"    {n}: ours {e:.2e}, reference fp32 {e32:.2e}, bar {max(2.5e-2 if name != 'cfg1' else 1e-1, 8 * e32):.2e}")
    worst_t = max(e for _, e, _, _ in trows)
    assert trows[0][0] < 1.0, trows[0]
    print(f"{name}: worst grad-norm err {worst:.2e}, worst grad-tensor err {worst_t:.2e}")
    sd = m.state_dict()
    for key in gold.files:
        if key.startswith("buf/"):
            assert _rel(sd[key[4:]].cpu().numpy(), gold[key]) < 1e-4, key


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_eval_features_match_reference(cuda_device, name):
    """eval-mode (running-stat BN) 512-d features as used by the SK feature sweep (src/sk_utils.py:185-211); the
    train step of the golden run is replayed first so the running statistics are the same."""
    from selavi_b200 import model as sv_model
    B, T, HW, ST, K, hc = CONFIGS[name]
    gold = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    video, spec, labels = make_inputs(name)
    m = build(sv_model.load_model, name).to(cuda_device).train()
    v, s = torch.from_numpy(video).to(cuda_device), torch.from_numpy(spec).to(cuda_device)
    with torch.no_grad():
        m(v, s)
        m.eval()
        m.return_features = True
        fv, fa = m(v, s)
    assert tuple(fv.shape) == (B, 512) and tuple(fa.shape) == (B, 512)
    ev, ea = _rel(fv.cpu().numpy(), gold["eval_feat_v"]), _rel(fa.cpu().numpy(), gold["eval_feat_a"])
    print(f"{name}: eval feature rel err video {ev:.2e} audio {ea:.2e}")
    assert ev < 1e-3 and ea < 1e-3


def test_benchmark_shape_streams_are_bit_reproducible(cuda_device):
    """configs[1] shapes (batch 16, 3x32x112x112 + 1x257x200, K=309, 10 heads), where the chunked BN-statistics reduction
    (nchunk > 1) runs on the video stream, the audio stream and the weight-gradient side streams at the same time: the
    step must be bit-identical run to run and with the audio / weight-gradient streams switched off (the
    per-device reduction scratch used to be shared between streams)."""
    from selavi_b200 import engine, model as sv_model
    from selavi_b200.utils import get_loss
    name = "big_cfg2"
    B, T, HW, ST, K, hc = CONFIGS[name]
    video, spec, labels = make_inputs(name)
    v, s, lab = (torch.from_numpy(x).to(cuda_device) for x in (video, spec, labels))

    def run(audio_stream, wgrad_stream):
        old = engine.AUDIO_STREAM, engine.WGRAD_STREAM
        engine.AUDIO_STREAM, engine.WGRAD_STREAM = audio_stream, wgrad_stream
        try:
            m = build(sv_model.load_model, name).to(cuda_device).train()
            fv, fa = m(v, s)
            loss = 0.5 * get_loss(fv, lab, headcount=hc) + 0.5 * get_loss(fa, lab, headcount=hc)
            loss.backward()
            torch.cuda.synchronize()
            return float(loss), {n: p.grad.clone() for n, p in m.named_parameters()}, {n: b.clone() for n, b in m.named_buffers()}
        finally:
            engine.AUDIO_STREAM, engine.WGRAD_STREAM = old

    l0, g0, b0 = run(True, True)
    for cfg in [(True, True), (False, True), (False, False)]:
        l1, g1, b1 = run(*cfg)
        assert l1 == l0, (cfg, l0, l1)
        for n in g0:
            assert torch.equal(g0[n], g1[n]), (cfg, n)
        for n in b0:
            assert torch.equal(b0[n], b1[n]), (cfg, n)


def test_ce_rejects_out_of_range_labels(cuda_device):
    """a label outside [0, K) poisons the loss and its gradient row with NaN (torch raises a device assert there)"""
    from selavi_b200.utils import get_loss
    x = torch.randn(4, 7, device=cuda_device, requires_grad=True)
    t = torch.tensor([0, 6, 7, 1], device=cuda_device)
    loss = get_loss(x, t)
    loss.backward()
    assert torch.isnan(loss)
    assert torch.isnan(x.grad[2]).all() and torch.isfinite(x.grad[[0, 1, 3]]).all()


def test_sgd_first_step_is_per_parameter(cuda_device):
    """a parameter that receives its first gradient at a later step must not reset the other parameters' momentum"""
    from selavi_b200.optim import SGD
    g = torch.Generator(device=cuda_device).manual_seed(4)
    ps = [torch.randn(n, device=cuda_device, generator=g) for n in (33, 1000)]
    a = [torch.nn.Parameter(p.clone()) for p in ps]
    b = [torch.nn.Parameter(p.clone()) for p in ps]
    oa = SGD(a, lr=0.05, momentum=0.9, weight_decay=1e-4)
    ob = torch.optim.SGD(b, lr=0.05, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        for i, (x, y) in enumerate(zip(a, b)):
            if i == 1 and step == 0:
                x.grad = y.grad = None          # second tensor: no gradient on the first step
                continue
            gr = torch.randn(x.shape, device=cuda_device, generator=g)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)


def test_sgd_matches_torch(cuda_device):
    from selavi_b200.optim import SGD
    g = torch.Generator(device=cuda_device).manual_seed(1)
    ps = [torch.randn(n, device=cuda_device, generator=g) for n in (5, 1000, 70001)]
    a = [torch.nn.Parameter(p.clone()) for p in ps]
    b = [torch.nn.Parameter(p.clone()) for p in ps]
    oa = SGD(a, lr=0.05, momentum=0.9, weight_decay=1e-4)
    ob = torch.optim.SGD(b, lr=0.05, momentum=0.9, weight_decay=1e-4)
    for _ in range(3):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, device=cuda_device, generator=g)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)


def test_heads_linear_and_dropout_paths(cuda_device):
    """use_mlp=False heads and train-mode dropout masks: gradients checked against torch autograd on the same masks."""
    from selavi_b200 import engine, model as sv_model
    torch.manual_seed(3)
    heads = [sv_model.MLPv2(512, 28).to(cuda_device).train() for _ in range(3)]
    x = torch.randn(6, 512, device=cuda_device, requires_grad=True)
    run = engine._HeadsRun(heads, True)
    logits, st = run.forward(x.detach(), save=True)
    dl = torch.randn_like(logits)
    dx, grads = run.backward(st, dl)
    # torch reference with the same masks
    xr = x.detach().double().requires_grad_(True)
    tot = 0
    for h, head in enumerate(heads):
        seq = head.block_forward
        w1, w2, b2 = seq[2].weight.double(), seq[8].weight.double(), seq[8].bias.double()
        z1 = (xr * st["m1"][h].double()) @ w1.t()
        mu, var = z1.mean(0), z1.var(0, unbiased=False)
        y1 = (z1 - mu) / torch.sqrt(var + 1e-5) * seq[4].weight.double() + seq[4].bias.double()
        a1 = torch.relu(y1) * st["m2"][h].double()
        lg = a1 @ w2.t() + b2
        torch.testing.assert_close(logits[h].double(), lg, rtol=1e-4, atol=1e-4)
        tot = tot + (lg * dl[h].double()).sum()
    ref_dx, = torch.autograd.grad(tot, xr)
    torch.testing.assert_close(dx.double(), ref_dx, rtol=1e-3, atol=1e-4)
    lin = [sv_model.LinearHead(512, 16).to(cuda_device) for _ in range(2)]
    outs = engine.heads_forward(lin, x)
    for o, l in zip(outs, lin):
        torch.testing.assert_close(o, torch.nn.functional.linear(x, l.weight, l.bias), rtol=1e-4, atol=1e-4)
    sum(o.sum() for o in outs).backward()
    assert x.grad is not None and lin[0].weight.grad is not None


def test_ce_loss_matches_torch(cuda_device):
    from selavi_b200.utils import get_loss
    g = torch.Generator(device=cuda_device).manual_seed(2)
    acts = [torch.randn(16, 309, device=cuda_device, generator=g, requires_grad=True) for _ in range(10)]
    tg = torch.randint(0, 309, (16, 10), device=cuda_device, generator=g)
    loss = get_loss(acts, tg, headcount=10)
    ref = torch.stack([torch.nn.functional.cross_entropy(a.detach().double(), tg[:, h]) for h, a in enumerate(acts)]).mean()
    assert abs(float(loss) - float(ref)) < 1e-5
    (loss * 2).backward()
    a0 = acts[3].detach().double().requires_grad_(True)
    (torch.nn.functional.cross_entropy(a0, tg[:, 3]) / 10 * 2).backward()
    torch.testing.assert_close(acts[3].grad.double(), a0.grad, rtol=1e-4, atol=1e-7)
    one = torch.randn(16, 28, device=cuda_device, requires_grad=True)
    t1 = torch.randint(0, 28, (16,), device=cuda_device)
    assert abs(float(get_loss(one, t1)) - float(torch.nn.functional.cross_entropy(one.detach(), t1))) < 1e-5
