"""Row-sharded label assignment (src/sk_utils.py:137-356 mirror): index bookkeeping on CPU with gloo (world_size 2),
end-to-end sweep on the GPU against the oracle solver."""
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from selavi_b200.sk_utils import assemble_labels, shard_range


def test_shard_range_drops_remainder_like_reference():
    N, ws = 3331, 8                                  # reference: local = N // world_size (src/sk_utils.py:157)
    covered = []
    for r in range(ws):
        lo, hi = shard_range(N, ws, r)
        assert hi - lo == N // ws
        covered += list(range(lo, hi))
    assert covered == list(range((N // ws) * ws))    # the last N % ws rows are never visited and keep label 0


def _worker(rank, world, port, N, hc, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N, world, rank)
    g = torch.Generator().manual_seed(rank)
    idx_local = torch.arange(lo, hi)[torch.randperm(hi - lo, generator=g)]      # arrival order of the sampler
    L = torch.zeros((N, hc), dtype=torch.long)
    for head in range(hc):
        lab_local = (idx_local * 7 + head) % 13                                   # stands in for the solver output
        gi = [torch.empty_like(idx_local) for _ in range(world)]
        gl = [torch.empty_like(lab_local) for _ in range(world)]
        dist.all_gather(gi, idx_local)
        dist.all_gather(gl, lab_local)
        assemble_labels(L, torch.cat(gi), torch.cat(gl), head)
    if rank == 0:
        torch.save(L, out)
    ok = torch.tensor([int(L.sum())])
    dist.all_reduce(ok)                                                          # every rank holds the same L
    assert int(ok) == world * int(L.sum())
    dist.destroy_process_group()


def test_label_assembly_two_ranks_gloo(tmp_path):
    N, hc, world = 1001, 3, 2
    out = str(tmp_path / "L.pt")
    mp.spawn(_worker, args=(world, 29653, N, hc, out), nprocs=world, join=True)
    L = torch.load(out)
    idx = torch.arange(N)
    for head in range(hc):
        exp = (idx * 7 + head) % 13
        exp[(N // world) * world:] = 0                                           # dropped remainder stays 0
        assert torch.equal(L[:, head], exp)


class _Clips(torch.utils.data.Dataset):
    def __init__(self, n):
        rng = np.random.default_rng(5)
        self.v = torch.from_numpy(rng.standard_normal((n, 3, 4, 32, 32)).astype(np.float32) * np.linspace(0.5, 2, n, dtype=np.float32).reshape(n, 1, 1, 1, 1))
        self.a = torch.from_numpy((rng.standard_normal((n, 1, 65, 40)) * 17.89 + 1.93).astype(np.float32))

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i], self.a[i], 0, i, i


@pytest.mark.gpu
def test_sweep_single_gpu_matches_oracle(cuda_device):
    from oracle.sk_oracle import optimize_L_sk
    from selavi_b200 import model as sv_model
    from selavi_b200.sk_utils import get_cluster_assignments_gpu, softmax_product
    torch.manual_seed(31)
    hc, K, N = 2, 8, 96
    m = sv_model.load_model(use_mlp=True, headcount=hc, num_classes=K, norm_feat=False).to(cuda_device)
    ds = _Clips(N)
    args = types.SimpleNamespace(world_size=1, rank=0, workers=0, ind_groups=1, headcount=hc, match=False, distribution="default",
                                 dist=None, diff_dist_every=False, diff_dist_per_head=True, gauss_sd=0.1, lamb=20.0,
                                 dump_path="")
    np.random.seed(0)
    L = get_cluster_assignments_gpu(args, ds, m, logger=None)
    assert tuple(L.shape) == (N, hc) and m.training and m.return_features is False
    # reference data flow on the same features: eval features -> heads -> f64 softmax product -> SK -> L[idx, head]
    m.eval()
    m.return_features = True
    with torch.no_grad():
        fv, fa = m(ds.v.to(cuda_device), ds.a.to(cuda_device))
        for head in range(hc):
            lv = getattr(m, f"mlp_v{head}").forward(fv)
            la = getattr(m, f"mlp_a{head}").forward(fa)
            PS = softmax_product(lv, la)
            ref = torch.softmax(lv.double(), 1) * torch.softmax(la.double(), 1)
            torch.testing.assert_close(PS, ref, rtol=1e-12, atol=0)
            ora = optimize_L_sk(PS.cpu().numpy())
            assert np.array_equal(L[:, head].cpu().numpy(), ora["labels"])


@pytest.mark.gpu
def test_match_order_matches_oracle(cuda_device):
    """match_order (K x K L1 cost matrix + host hill-climb on it) picks the same permutation as the reference's
    column-swapping search for the same np.random stream, and permutes the Linear rows accordingly."""
    from oracle.sk_oracle import match_order_oracle
    from selavi_b200.sk_utils import match_order
    rng = np.random.default_rng(3)
    N, K = 400, 12
    true_perm = rng.permutation(K)
    lv = rng.standard_normal((N, K)) * 3
    la = lv[:, true_perm] + 0.05 * rng.standard_normal((N, K))            # audio head = permuted video head + noise

    def sm(x):
        e = np.exp(x - x.max(1, keepdims=True))
        return e / e.sum(1, keepdims=True)

    np.random.seed(11)
    ref = match_order_oracle(sm(lv), sm(la))
    lin = torch.nn.Linear(5, K).to(cuda_device)
    w0, b0 = lin.weight.data.clone(), lin.bias.data.clone()
    args = types.SimpleNamespace(rank=0)
    np.random.seed(11)
    fin = match_order(args, torch.from_numpy(lv).float().to(cuda_device), torch.from_numpy(la).float().to(cuda_device), lin,
                      logits=True)
    assert np.array_equal(fin.cpu().numpy(), ref)
    assert torch.equal(lin.weight.data, w0[fin]) and torch.equal(lin.bias.data, b0[fin])
    # the search must actually have aligned the heads: emb2[:, perm] ~ emb1
    assert np.abs(sm(lv) - sm(la)[:, ref]).sum() < 0.2 * np.abs(sm(lv) - sm(la)).sum()
