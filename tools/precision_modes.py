"""Precision experiments on the GPU box (informational, feeds DESIGN.md):
  1. fast mode (engine.PASSES = 1) logit error vs parity mode, eval and train mode, fresh weights (mini_cfg2 and big_cfg2)
  2. big_cfg2 gradient-norm errors vs float64 for SELAVI_BWD = bf16x3 (default) and tf32x3
usage: python tools/precision_modes.py [fast] [bwd]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gen_golden_model import CONFIGS, build, make_inputs  # noqa: E402
from selavi_b200 import engine, model as sv_model  # noqa: E402
from selavi_b200.utils import get_loss  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1:] or ["fast", "bwd"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


if "fast" in what:
    for name in ("mini_cfg2", "big_cfg2"):
        B, T, HW, ST, K, hc = CONFIGS[name]
        video, spec, _ = make_inputs(name)
        v, s = torch.from_numpy(video).to(dev), torch.from_numpy(spec).to(dev)
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"model_{name}.npz"))
        for mode in ("train", "eval"):
            outs = {}
            for passes in (3, 1):
                engine.PASSES = passes
                m = build(sv_model.load_model, name).to(dev)
                m.train() if mode == "train" else m.eval()
                with torch.no_grad():
                    fv, fa = m(v, s)
                outs[passes] = (torch.stack(list(fv)), torch.stack(list(fa)))
            engine.PASSES = 3
            line = f"{name} {mode}: fast-vs-parity logits video {rel(outs[1][0], outs[3][0]):.2e} audio {rel(outs[1][1], outs[3][1]):.2e}"
            if mode == "train":
                g64 = torch.from_numpy(gold["logits_v64"]).to(dev)
                line += f" | vs float64: parity {rel(outs[3][0], g64):.2e} fast {rel(outs[1][0], g64):.2e}"
            print(line, flush=True)

if "bwd" in what:
    name = "big_cfg2"
    B, T, HW, ST, K, hc = CONFIGS[name]
    video, spec, labels = make_inputs(name)
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"model_{name}.npz"))
    v, s, lab = (torch.from_numpy(x).to(dev) for x in (video, spec, labels))
    for bwd in ("bf16x3", "tf32x3"):
        engine.BWD = bwd
        m = build(sv_model.load_model, name).to(dev).train()
        fv, fa = m(v, s)
        loss = 0.5 * get_loss(fv, lab, headcount=hc) + 0.5 * get_loss(fa, lab, headcount=hc)
        loss.backward()
        grads = {n: p.grad for n, p in m.named_parameters()}
        rows = []
        for n, n32, n64 in zip(gold["grad_names"], gold["grad_norms"], gold["grad_norms64"]):
            e = abs(float(grads[str(n)].norm()) - n64) / max(n64, 1e-12)
            rows.append((e, abs(n32 - n64) / max(n64, 1e-12), str(n)))
        rows.sort(reverse=True)
        print(f"{name} backward {bwd}: gradient-norm error vs float64 (ours, reference fp32):")
        for e, e32, n in rows[:8]:
            print(f"    {e:.2e} {e32:.2e} {n}")
        es = np.array([r[0] for r in rows]); e32s = np.array([r[1] for r in rows])
        print(f"    median ours {np.median(es):.2e} ref32 {np.median(e32s):.2e}; ours > 5e-3: {(es > 5e-3).sum()} of {len(es)}; "
              f"ours > 8x ref32 and > 5e-3: {((es > 8 * e32s) & (es > 5e-3)).sum()}")
        trows = []
        for key in gold.files:
            if key.startswith("grad64/"):
                n = key[len("grad64/"):]
                g64 = torch.from_numpy(gold[key]).to(dev)
                trows.append((rel(grads[n], g64), rel(torch.from_numpy(gold["grad/" + n]).to(dev), g64), n))
        trows.sort(reverse=True)
        print("    full tensors (ours, reference fp32):")
        for e, e32, n in trows[:6]:
            print(f"    {e:.2e} {e32:.2e} {n}")
        del m, grads
    engine.BWD = "bf16x3"
