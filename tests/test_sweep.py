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


@pytest.mark.gpu
@pytest.mark.parametrize("N,K", [(7, 12), (400, 28), (5000, 309), (1531, 400)])
def test_l1_cost_matrix_kernel(cuda_device, N, K):
    """csrc/match.cu: C[i,j] = sum_n |P1[n,i] - P2[n,j]| and the float64 softmax feeding it, against numpy."""
    from oracle.sk_oracle import softmax64 as sm
    from selavi_b200.sk_utils import l1_cost_matrix, softmax64
    rng = np.random.default_rng(N + K)
    lv = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    la = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    Pv, Pa = softmax64(torch.from_numpy(lv).to(cuda_device)), softmax64(torch.from_numpy(la).to(cuda_device))
    np.testing.assert_allclose(Pv.cpu().numpy(), sm(lv), rtol=1e-13, atol=0)
    C = l1_cost_matrix(Pv, Pa).cpu().numpy()
    ref = np.abs(sm(lv)[:, :, None] - sm(la)[:, None, :]).sum(0) if N * K * K < 5e7 else \
        np.stack([np.abs(sm(lv)[:, i:i + 1] - sm(la)).sum(0) for i in range(K)])
    np.testing.assert_allclose(C, ref, rtol=1e-12, atol=1e-15)
    C2 = l1_cost_matrix(Pv, Pa).cpu().numpy()
    assert np.array_equal(C, C2)                     # fixed-order split-N reduce: bit-reproducible


@pytest.mark.gpu
def test_match_order_k309_matches_oracle(cuda_device):
    """cfg-2/3 size of the head alignment (K = 309): same permutation as the reference's search for the same np.random
    stream, although every cost comes from the precomputed K x K matrix instead of per-step column sums."""
    from oracle.sk_oracle import match_order_oracle, softmax64 as sm
    from selavi_b200.sk_utils import match_order
    rng = np.random.default_rng(9)
    N, K = 2000, 309
    true_perm = rng.permutation(K)
    lv = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    la = (lv[:, true_perm] + 0.05 * rng.standard_normal((N, K))).astype(np.float32)
    np.random.seed(5)
    ref = match_order_oracle(sm(lv), sm(la), steps=6000)
    lin = torch.nn.Linear(5, K).to(cuda_device)
    w0 = lin.weight.data.clone()
    np.random.seed(5)
    fin = match_order(types.SimpleNamespace(rank=0), torch.from_numpy(lv).to(cuda_device), torch.from_numpy(la).to(cuda_device),
                      lin, steps=6000, logits=True)
    assert np.array_equal(fin.cpu().numpy(), ref)
    assert torch.equal(lin.weight.data, w0[fin])


def _sweep_args(**kw):
    base = dict(world_size=1, rank=0, workers=0, ind_groups=1, headcount=1, match=False, distribution="default", dist=None,
                diff_dist_every=False, diff_dist_per_head=True, gauss_sd=0.1, lamb=20.0, dump_path="")
    base.update(kw)
    return types.SimpleNamespace(**base)


def _warm_bn(m, ds, dev, passes=25):
    """Train-mode no-grad forwards (the reference's `warmup_batchnorm`, utils.py:389-418, run by main.py before the first SK
    call): with the constructor's running statistics (mean 0, var 1) the eval-mode features of an untrained tower are
    nearly collinear across clips and the head outputs differ mostly by per-column offsets — about 40 % of the swap
    deltas of `match_order` are then EXACT ties in exact arithmetic and their sign is rounding noise, in the reference
    as much as here.  Warmed up, the smallest |delta| seen is 1e-5 of its scale."""
    m.train()
    n = len(ds)
    with torch.no_grad():
        for it in range(passes):
            lo = (it * 32) % max(1, n - 40)
            m(ds.v[lo:lo + 40].to(dev), ds.a[lo:lo + 40].to(dev))


def _head_logits(m, ds, dev, hc):
    """eval-mode outputs of every head on the whole dataset, dataset-index order (inputs of the bookkeeping oracle)"""
    m.eval()
    with torch.no_grad():
        if hc == 1:
            lv, la = m(ds.v.to(dev), ds.a.to(dev))
            out = [lv.cpu().numpy()], [la.cpu().numpy()]
        else:
            m.return_features = True
            fv, fa = m(ds.v.to(dev), ds.a.to(dev))
            m.return_features = False
            out = ([getattr(m, f"mlp_v{h}").forward(fv).cpu().numpy() for h in range(hc)],
                   [getattr(m, f"mlp_a{h}").forward(fa).cpu().numpy() for h in range(hc)])
    m.train()
    return out


@pytest.mark.gpu
def test_sweep_cfg4_match_ind_groups(cuda_device):
    """BASELINE.json configs[3] (cfg-4) flow of `get_cluster_assignments_gpu`: K = 28, 10 heads, `match` on the first SK
    call, ind_groups = 2 (two sweeps, heads dealt `order_heads[g::2]`), per-head Gaussian marginals, N = 3328 (AVE size).
    Checked against the CPU bookkeeping oracle (oracle/sk_oracle.py:cluster_assignments_oracle, src/sk_utils.py:183-327)
    fed with the same head outputs: labels of every head, the permutation applied to every audio head, model state."""
    from oracle.sk_oracle import cluster_assignments_oracle
    from selavi_b200 import model as sv_model
    from selavi_b200.sk_utils import get_cluster_assignments_gpu
    torch.manual_seed(31)
    hc, K, N = 10, 28, 3328
    m = sv_model.load_model(use_mlp=True, headcount=hc, num_classes=K, norm_feat=False).to(cuda_device)
    ds = _Clips(N)
    _warm_bn(m, ds, cuda_device)
    lv, la = _head_logits(m, ds, cuda_device, hc)
    rng = np.random.default_rng(2)
    kd = [(rng.standard_normal(K) * 0.1 + 1) * N / K for _ in range(hc)]
    args = _sweep_args(ind_groups=2, headcount=hc, match=True, distribution="gauss",
                       dist=[torch.from_numpy(k.copy()).view(K, 1).to(cuda_device) for k in kd])
    w_before = [list(getattr(m, f"mlp_a{h}").modules())[-1].weight.data.clone() for h in range(hc)]
    np.random.seed(7)
    L = get_cluster_assignments_gpu(args, ds, m, logger=None, iter_num=0)
    np.random.seed(7)
    L_ref, perms, order = cluster_assignments_oracle(lv, la, N, 1, 2, True, kdists=kd)
    assert sorted(perms) == list(range(hc)) and sorted(order) == list(range(hc))
    for h in range(hc):
        lin = list(getattr(m, f"mlp_a{h}").modules())[-1]
        assert torch.equal(lin.weight.data, w_before[h][torch.from_numpy(perms[h]).to(cuda_device)]), h
        assert np.array_equal(L[:, h].cpu().numpy(), L_ref[:, h]), (h, int((L[:, h].cpu().numpy() != L_ref[:, h]).sum()))
    assert m.training and m.return_features is False
    # a later SK call (iter_num > 0) must not re-align the heads
    np.random.seed(8)
    w_mid = [list(getattr(m, f"mlp_a{h}").modules())[-1].weight.data.clone() for h in range(hc)]
    get_cluster_assignments_gpu(args, ds, m, logger=None, iter_num=5)
    for h in range(hc):
        assert torch.equal(list(getattr(m, f"mlp_a{h}").modules())[-1].weight.data, w_mid[h])


@pytest.mark.gpu
@pytest.mark.parametrize("match", [False, True])
def test_sweep_single_head_branch(cuda_device, match):
    """headcount == 1 branch (src/sk_utils.py:206-211,301-303): the model's own outputs are soft-maxed and multiplied;
    with `match` the audio head is permuted but the assignment still uses the outputs gathered before it."""
    from oracle.sk_oracle import cluster_assignments_oracle
    from selavi_b200 import model as sv_model
    from selavi_b200.sk_utils import get_cluster_assignments_gpu
    torch.manual_seed(31)
    K, N = 28, 203                       # N % world_size == 0 here; the remainder rule is covered by the gloo test
    m = sv_model.load_model(use_mlp=True, headcount=1, num_classes=K, norm_feat=False).to(cuda_device)
    ds = _Clips(N)
    _warm_bn(m, ds, cuda_device)
    lv, la = _head_logits(m, ds, cuda_device, 1)
    args = _sweep_args(match=match)
    w0 = list(m.mlp_a.modules())[-1].weight.data.clone()
    np.random.seed(3)
    L = get_cluster_assignments_gpu(args, ds, m, logger=None, iter_num=0)
    np.random.seed(3)
    L_ref, perms, _ = cluster_assignments_oracle(lv, la, N, 1, 1, match)
    assert tuple(L.shape) == (N, 1)
    assert np.array_equal(L[:, 0].cpu().numpy(), L_ref[:, 0])
    if match:
        assert torch.equal(list(m.mlp_a.modules())[-1].weight.data, w0[torch.from_numpy(perms[0]).to(cuda_device)])
    assert m.training and m.return_features is False


@pytest.mark.gpu
def test_sweep_against_cpu_oracle_towers(cuda_device):
    """End to end against the CPU oracle INCLUDING the towers (oracle/model_oracle.py in eval mode -> heads -> softmax ->
    SK): the GPU features carry ~1e-4 relative rounding differences, so rows whose two best clusters are within that
    margin may flip; at least 97 % of the labels must agree (measured: see the printed fraction)."""
    from oracle.model_oracle import OracleAVModel
    from oracle.sk_oracle import cluster_assignments_oracle
    from selavi_b200 import model as sv_model
    from selavi_b200.sk_utils import get_cluster_assignments_gpu
    hc, K, N = 2, 8, 160
    torch.manual_seed(31)
    m = sv_model.load_model(use_mlp=True, headcount=hc, num_classes=K, norm_feat=False)
    o = OracleAVModel(hc, K)
    o.load_state_dict(m.state_dict())
    ds = _Clips(N)
    o.eval()
    with torch.no_grad():
        ov, oa = o(ds.v, ds.a)
    m = m.to(cuda_device)
    np.random.seed(1)
    L = get_cluster_assignments_gpu(_sweep_args(headcount=hc), ds, m, logger=None)
    np.random.seed(1)
    L_ref, _, _ = cluster_assignments_oracle([x.numpy() for x in ov], [x.numpy() for x in oa], N, 1, 1, False)
    agree = float((L.cpu().numpy() == L_ref).mean())
    print(f"sweep vs full CPU oracle (towers included): {agree * 100:.1f} % of {N * hc} labels agree")
    assert agree >= 0.97


def _ddp_ignore_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from selavi_b200 import model as sv_model
    torch.manual_seed(31 + rank)             # different initial weights per rank on purpose
    m = sv_model.load_model(use_mlp=True, headcount=2, num_classes=8, norm_feat=False)
    assert not hasattr(torch.nn.Module(), "_ddp_params_and_buffers_to_ignore")
    plain = hasattr(m, "_ddp_params_and_buffers_to_ignore")          # plain BatchNorm: DDP keeps its stock behaviour
    m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
    names = set(m._ddp_params_and_buffers_to_ignore)
    # what DistributedDataParallel.__init__ does with the attribute (torch/nn/parallel/distributed.py:723-736); DDP itself
    # refuses SyncBatchNorm modules on the CPU, the real construction is exercised by tests/test_multigpu.py
    assert hasattr(m, "_ddp_params_and_buffers_to_ignore")
    # any reader other than DDP's constructor gets the names without side effects (no collective, towers untouched)
    assert m._ddp_params_and_buffers_to_ignore and "_sv_runner" not in m.video_network.base.__dict__
    # the same access made from a frame of module torch.nn.parallel.distributed, as DistributedDataParallel.__init__ makes it
    import types
    code = compile("def read(module):\n    return module._ddp_params_and_buffers_to_ignore\n", "<ddp>", "exec")
    read = types.FunctionType(code.co_consts[0], {"__name__": "torch.nn.parallel.distributed"})
    ignore = set(read(m))
    reduced_by_ddp = [n for n, _ in m.named_parameters() if n not in ignore]
    # the ignored tensors were synchronised to rank 0's values here (DDP's own construction-time broadcast skips them)
    probe = torch.cat([m.video_network.base.stem[0].weight.detach().reshape(-1)[:64],
                       m.audio_network.base.layer4[0].conv2.weight.detach().reshape(-1)[:64]])
    gathered = [torch.empty_like(probe) for _ in range(world)]
    dist.all_gather(gathered, probe)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        torch.save(dict(plain=plain, names=sorted(names), reduced=reduced_by_ddp, buffers=[n for n, _ in m.named_buffers()],
                        tower=[n for n, _ in m.named_parameters() if n.startswith(("video_network", "audio_network"))],
                        own=(m.video_network.base.__dict__["_sv_runner"].own_allreduce, m.audio_network.base.__dict__["_sv_runner"].own_allreduce)),
                   out)
    dist.destroy_process_group()


def test_ddp_ignores_tower_params_and_bn_buffers_gloo(tmp_path):
    """engine.DDP_BYPASS: a SyncBatchNorm-converted model tells DistributedDataParallel to leave the towers' parameters and
    all BatchNorm buffers alone (the engine averages those gradients itself, overlapped with backward); the heads stay
    with DDP; a model with plain BatchNorm does not expose the attribute at all."""
    out = str(tmp_path / "ignore.pt")
    mp.spawn(_ddp_ignore_worker, args=(2, 29671, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["plain"] is False
    assert set(r["buffers"]) <= set(r["names"]) and set(r["tower"]) <= set(r["names"])
    assert r["reduced"] and all(n.startswith("mlp_") for n in r["reduced"])
    assert r["own"] == (True, True)
