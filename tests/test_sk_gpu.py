"""GPU parity of the Sinkhorn-Knopp kernel (through the C ABI) against the oracle and the reference golden vectors."""
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_golden_sk import CASES, kdist_for  # noqa: E402
from oracle.sk_oracle import optimize_L_sk, synth_PS, top2_margin  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sk_cases.npz"))


class _Log:
    def __init__(self):
        self.lines = []

    def info(self, s):
        self.lines.append(str(s))


def _args(distribution="default", per_head=True, headcount=1, lamb=20.0, dist=None):
    return types.SimpleNamespace(distribution=distribution, diff_dist_every=False, diff_dist_per_head=per_head,
                                 gauss_sd=0.1, headcount=headcount, lamb=lamb, rank=0, dist=dist)


def _assert_labels(labels, ref_labels, PS, lamb, alpha, beta):
    labels = np.asarray(labels)
    bad = np.nonzero(labels != ref_labels)[0]
    if len(bad):
        # a genuine mismatch has a visible gap between best and second best; ulp-level ties do not
        margin = top2_margin(PS[bad], lamb, alpha, beta[bad])
        assert np.all(margin < 1e-9), f"{len(bad)} label mismatches, max margin {margin.max()}"
    assert len(bad) == 0, f"{len(bad)} tie-level label mismatches"


@pytest.mark.parametrize("name", sorted(CASES))
def test_sk_matches_reference_golden(cuda_device, name):
    from selavi_b200.sk_utils import optimize_L_sk_gpu
    N, K, scale, seed, distribution, per_head, headcount, hc, lamb = CASES[name]
    PS_np = synth_PS(N, K, scale, seed)
    kd = kdist_for(name)
    dist = None
    if kd is not None:
        dist = ([torch.from_numpy(k.copy()).to(cuda_device) for k in kd] if per_head
                else torch.from_numpy(kd.copy()).to(cuda_device))
    args = _args(distribution, per_head, headcount, lamb, dist)
    log = _Log()
    cost, L = optimize_L_sk_gpu(args, torch.from_numpy(PS_np).to(cuda_device), hc, logger=log)
    ora = optimize_L_sk(PS_np, lamb=lamb, kdist=None if kd is None else (kd[hc] if per_head else kd))
    assert f"step : {int(GOLD[name + '/iters'])}" in "\n".join(log.lines)
    _assert_labels(L.cpu().numpy(), GOLD[name + "/labels"].astype(np.int64), PS_np, lamb, ora["alpha"], ora["beta"])
    assert abs(cost - float(GOLD[name + "/cost"])) <= 1e-9 * abs(float(GOLD[name + "/cost"]))
    if kd is not None:
        after = args.dist[hc] if per_head else args.dist
        np.testing.assert_array_equal(after.cpu().numpy().reshape(-1), GOLD[name + "/kdist_after"])


@pytest.mark.parametrize("N,K,scale", [(20000, 309, 1.0), (20001, 309, 4.0), (7, 28, 1.0), (2, 5, 1.0), (30000, 400, 1.0),
                                       (5000, 512, 1.0), (4097, 33, 2.0)])
def test_sk_matches_oracle(cuda_device, N, K, scale):
    from selavi_b200.sk_utils import SKWorkspace, sk_solve_raw
    PS_np = synth_PS(N, K, scale, seed=N + K)
    ora = optimize_L_sk(PS_np)
    PS = torch.from_numpy(PS_np).to(cuda_device)
    ws = SKWorkspace(K, N, cuda_device)
    sk_solve_raw(PS, N, 20.0, None, ws)
    assert int(ws.iters.item()) == ora["iters"]
    np.testing.assert_allclose(ws.alpha.cpu().numpy(), ora["alpha"], rtol=1e-10)
    np.testing.assert_allclose(ws.beta.cpu().numpy(), ora["beta"], rtol=1e-10)
    np.testing.assert_allclose(PS.cpu().numpy(), PS_np ** 10.0, rtol=1e-13)      # PS consumed: pow_ in place
    _assert_labels(ws.labels.cpu().numpy(), ora["labels"], PS_np, 20.0, ora["alpha"], ora["beta"])
    cost = -(1.0 / 20.0) * float(ws.cost.item()) / N
    assert abs(cost - ora["cost"]) <= 1e-9 * abs(ora["cost"])
    assert abs(float(ws.err.item()) - ora["err"]) <= 1e-6 * ora["err"]


def test_sk_fixed_iterations_and_continuation(cuda_device):
    """stop_on_converge=0 runs exactly max_iters; do_prep=0 continues from the workspace state (bench mode)."""
    from selavi_b200.sk_utils import SKWorkspace, sk_solve_raw
    N, K = 6000, 309
    PS_np = synth_PS(N, K, 1.0, 5)
    ora = optimize_L_sk(PS_np, max_iters=37, stop_on_converge=False)
    PS = torch.from_numpy(PS_np).to(cuda_device)
    ws = SKWorkspace(K, N, cuda_device)
    sk_solve_raw(PS, N, 20.0, None, ws, max_iters=20, stop_on_converge=False, do_final=False)
    assert int(ws.iters.item()) == 20
    sk_solve_raw(PS, N, 20.0, None, ws, max_iters=17, stop_on_converge=False, do_prep=False, do_final=True)
    assert int(ws.iters.item()) == 17
    np.testing.assert_allclose(ws.alpha.cpu().numpy(), ora["alpha"], rtol=1e-10)
    np.testing.assert_allclose(ws.beta.cpu().numpy(), ora["beta"], rtol=1e-10)
    assert np.array_equal(ws.labels.cpu().numpy(), ora["labels"])


def test_sk_full_size_properties(cuda_device):
    """cfg-5 size (N=200000, K=309): size-independent properties — the plan's row sums are 1/N, its column
    sums follow r within the stopping tolerance, labels are the row argmax, and two runs are bit-identical."""
    from selavi_b200.sk_utils import SKWorkspace, sk_solve_raw
    N, K = 200000, 309
    g = torch.Generator(device=cuda_device).manual_seed(0)
    a = torch.softmax(torch.randn(N, K, dtype=torch.float64, device=cuda_device, generator=g), 1)
    b = torch.softmax(torch.randn(N, K, dtype=torch.float64, device=cuda_device, generator=g), 1)
    PS0 = a * b
    outs = []
    for _ in range(2):
        PS = PS0.clone()
        ws = SKWorkspace(K, N, cuda_device)
        sk_solve_raw(PS, N, 20.0, None, ws)
        P = PS * ws.beta[:, None] * ws.alpha[None, :]
        assert 10 <= int(ws.iters.item()) < 2000 and float(ws.err.item()) <= 0.1
        torch.testing.assert_close(PS, PS0 ** 10, rtol=1e-13, atol=0)
        torch.testing.assert_close(P.sum(1), torch.full((N,), 1.0 / N, dtype=torch.float64, device=cuda_device),
                                   rtol=1e-12, atol=0)
        assert float((P.sum(0) * K - 1).abs().max()) < 0.05
        assert torch.equal(P.argmax(1), ws.labels)
        outs.append((ws.labels.clone(), ws.alpha.clone(), ws.beta.clone()))
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x, y)


def test_sk_rejects_bad_input(cuda_device):
    from selavi_b200 import _lib
    from selavi_b200.sk_utils import SKWorkspace, sk_solve_raw
    with pytest.raises(ValueError):
        sk_solve_raw(torch.zeros(4, 4, device=cuda_device), 4, 20.0, None, None)
    with pytest.raises(_lib.SelaviError):
        SKWorkspace(600, 10, cuda_device)
