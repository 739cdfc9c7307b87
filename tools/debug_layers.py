"""Per-layer comparison of the B200 towers (train mode, with tape) against the CPU oracle model: prints the relative
error of every BatchNorm output (pre-ReLU) in execution order.  Usage: python tools/debug_layers.py [cfg1|mini_cfg2]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gen_golden_model import CONFIGS, make_inputs  # noqa: E402
from oracle.model_oracle import OracleAVModel  # noqa: E402
from selavi_b200 import engine, model as sv_model, ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mini_cfg2"
B, T, HW, ST, K, hc = CONFIGS[name]
video, spec, labels = make_inputs(name)
dev = torch.device("cuda:0")
torch.manual_seed(31)
m = sv_model.load_model(use_mlp=True, headcount=hc, num_classes=K, norm_feat=False).to(dev).train()
torch.manual_seed(31)
o = OracleAVModel(hc, K).train()

for tower, x in (("video", video), ("audio", spec)):
    onet = getattr(o, tower + "_network").base
    mnet = getattr(m, tower + "_network").base
    outs = []
    hooks = []
    for n_, mod in onet.named_modules():
        if isinstance(mod, (torch.nn.BatchNorm3d, torch.nn.BatchNorm2d)):
            hooks.append(mod.register_forward_hook(lambda mod_, i, out, n_=n_: outs.append((n_, out.detach().clone()))))
    with torch.no_grad():
        ofeat = onet(torch.from_numpy(x))
    for h in hooks:
        h.remove()
    runner = engine.TowerRunner(mnet, tower)
    tape = []
    with torch.no_grad():
        feat = runner.forward(mnet, torch.from_numpy(x).to(dev), True, tape)
    recs = []
    for e in tape:
        if e[0] == "vstem":
            recs += [e[1], e[2]]
        elif e[0] == "astem":
            recs += [e[1]]
        elif e[0] == "block":
            recs += list(e[1]) + ([e[2]] if e[2] is not None else [])
    # oracle BN order inside a block with downsample: main path BNs then downsample BN (same as ours)
    print(f"--- {tower}: {len(recs)} conv+BN units, {len(outs)} oracle BN outputs")
    for rec, (n_, ref) in zip(recs, outs):
        z = ops.from_channels_last(rec.z, rec.geom.co).cpu()
        if ref.dim() == 4:
            z = z[:, :, 0]
        sc, sh = rec.scale[:rec.geom.co].cpu(), rec.shift[:rec.geom.co].cpu()
        shape = [1, -1] + [1] * (ref.dim() - 2)
        mine = z * sc.view(shape) + sh.view(shape)
        err = float((mine - ref).norm() / ref.norm())
        print(f"{n_:32s} {tuple(ref.shape)} rel={err:.3e}  invstd_max={float(rec.invstd.max()):.3e}")
    print(f"{tower} feature rel={float((feat.cpu() - ofeat.flatten(1)).norm() / ofeat.norm()):.3e}")
