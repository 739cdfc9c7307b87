"""The SK oracle (oracle/sk_oracle.py) is pinned against outputs of the UNMODIFIED reference solver
(tests/golden/sk_cases.npz, produced by tests/golden/gen_golden_sk.py from /root/reference)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_golden_sk import CASES, kdist_for  # noqa: E402
from oracle.sk_oracle import optimize_L_sk, synth_PS  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sk_cases.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    N, K, scale, seed, distribution, per_head, headcount, hc, lamb = CASES[name]
    PS = synth_PS(N, K, scale, seed)
    kd = kdist_for(name)
    kdist = None if kd is None else (kd[hc] if per_head else kd)
    out = optimize_L_sk(PS, lamb=lamb, kdist=kdist)
    assert out["iters"] == int(GOLD[name + "/iters"])
    assert np.array_equal(out["labels"], GOLD[name + "/labels"].astype(np.int64))
    assert abs(out["cost"] - float(GOLD[name + "/cost"])) <= 1e-9 * abs(float(GOLD[name + "/cost"]))
    assert abs(out["err"] - float(GOLD[name + "/err"])) <= 1e-6 * float(GOLD[name + "/err"])
    if kd is not None:
        np.testing.assert_array_equal(out["kdist"], GOLD[name + "/kdist_after"])


def test_oracle_invariants():
    # row sums of the scaled plan are 1/N, column sums follow r
    PS = synth_PS(600, 28, 1.0, 11)
    out = optimize_L_sk(PS, tol=1e-9, max_iters=500)
    P = (PS ** 10) * out["beta"][:, None] * out["alpha"][None, :]
    np.testing.assert_allclose(P.sum(1), 1.0 / 600, rtol=1e-9)
    np.testing.assert_allclose(P.sum(0), 1.0 / 28, rtol=1e-6)


def test_model_oracle_matches_reference_golden():
    """oracle/model_oracle.py (torchvision/torch.nn restatement of model.py + get_loss) reproduces the golden
    outputs of the real reference (tests/golden/model_cfg1.npz) bit for bit."""
    import torch
    from gen_golden_model import CONFIGS, make_inputs
    from oracle.model_oracle import OracleAVModel, oracle_get_loss
    name = "cfg1"
    B, T, HW, ST, K, hc = CONFIGS[name]
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"model_{name}.npz"))
    video, spec, labels = make_inputs(name)
    torch.manual_seed(31)
    m = OracleAVModel(hc, K)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    fv, fa = m(torch.from_numpy(video), torch.from_numpy(spec))
    lab = torch.from_numpy(labels)[:, 0]
    loss = 0.5 * oracle_get_loss(fv, lab) + 0.5 * oracle_get_loss(fa, lab)
    loss.backward()
    assert np.array_equal(fv.detach().numpy()[None], gold["logits_v"])
    assert np.array_equal(fa.detach().numpy()[None], gold["logits_a"])
    assert float(loss) == float(gold["loss"])
    norms = np.array([float(p.grad.norm()) for _, p in m.named_parameters()])
    np.testing.assert_allclose(norms, gold["grad_norms"], rtol=1e-6)
