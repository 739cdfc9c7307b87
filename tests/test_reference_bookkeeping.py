"""The bookkeeping oracle (oracle/sk_oracle.py:cluster_assignments_oracle, match_order_oracle) pinned against the
UNMODIFIED reference `get_cluster_assignments_gpu` / `match_order` / `optimize_L_sk_gpu` (src/sk_utils.py:137-467) run here
on the CPU: the reference source is loaded with its CUDA-only spellings neutralised at text level (oracle/ref_loader.py,
no arithmetic or control-flow change), under a single-process gloo group, on a torch CPU model (oracle/model_oracle.py,
itself pinned bit for bit against the reference model).  This is what the GPU sweep tests (tests/test_sweep.py,
tests/mgpu_worker.py) use as their checker."""
import copy
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist

from oracle import ref_loader
from oracle.model_oracle import OracleAVModel
from oracle.sk_oracle import cluster_assignments_oracle

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


class _Clips(torch.utils.data.Dataset):
    def __init__(self, n):
        r = np.random.default_rng(5)
        self.v = torch.from_numpy(r.standard_normal((n, 3, 4, 32, 32)).astype(np.float32) * np.linspace(0.5, 2, n, dtype=np.float32).reshape(n, 1, 1, 1, 1))
        self.a = torch.from_numpy((r.standard_normal((n, 1, 65, 40)) * 17.89 + 1.93).astype(np.float32))

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i], self.a[i], 0, np.int32(i), i        # int32 index: the reference gathers it into an IntTensor


class _DDP:
    """what the reference touches on the DistributedDataParallel wrapper"""

    def __init__(self, m):
        self.module = m

    def __call__(self, v, a):
        return self.module(v, a)

    def eval(self):
        self.module.eval()

    def train(self):
        self.module.train()


class _Log:
    def info(self, *_a, **_k):
        pass


@pytest.fixture(scope="module")
def gloo_group():
    os.environ["SIGNAL_RECEIVED"] = "False"
    dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:29688")
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("hc,K,match,ind_groups,distribution", [(3, 12, True, 2, "default"), (1, 12, True, 1, "default"),
                                                                (2, 8, False, 1, "gauss")])
def test_bookkeeping_oracle_matches_reference(gloo_group, hc, K, match, ind_groups, distribution):
    ref = ref_loader.load_sk_module("cpu", sweep=True)
    N = 120
    ds = _Clips(N)
    torch.manual_seed(31)
    m = OracleAVModel(hc, K)
    m.use_mlp = True
    m.train()
    with torch.no_grad():              # warmup_batchnorm (utils.py:389-418): see tests/test_sweep.py:_warm_bn
        for it in range(20):
            lo = (it * 16) % 80
            m(ds.v[lo:lo + 40], ds.a[lo:lo + 40])
    # inputs of the oracle: every head's outputs on the whole dataset, dataset order, BEFORE any head permutation
    m0 = copy.deepcopy(m).eval()
    with torch.no_grad():
        if hc == 1:
            lv, la = m0(ds.v, ds.a)
            lv, la = [lv.numpy()], [la.numpy()]
        else:
            m0.return_features = True
            fv, fa = m0(ds.v, ds.a)
            lv = [getattr(m0, f"mlp_v{h}")(fv).numpy() for h in range(hc)]
            la = [getattr(m0, f"mlp_a{h}")(fa).numpy() for h in range(hc)]
    heads_a = [m.mlp_a] if hc == 1 else [getattr(m, f"mlp_a{h}") for h in range(hc)]
    w0 = [list(h.modules())[-1].weight.data.clone() for h in heads_a]
    kd = None
    args = types.SimpleNamespace(world_size=1, rank=0, workers=0, ind_groups=ind_groups, headcount=hc, match=match,
                                 distribution=distribution, dist=None, diff_dist_every=False, diff_dist_per_head=True,
                                 gauss_sd=0.1, lamb=20, dump_path="")
    if distribution == "gauss":        # preset marginals (the reference would draw them with torch.randn on its device)
        rng = np.random.default_rng(2)
        kd = [(rng.standard_normal(K) * 0.1 + 1) * N / K for _ in range(hc)]
        args.dist = [torch.from_numpy(k.copy()).view(K, 1) for k in kd]
    np.random.seed(7)
    torch.manual_seed(7)
    with torch.no_grad():
        L_ref = ref.get_cluster_assignments_gpu(args, ds, _DDP(m), logger=_Log(), writer=None, group=None, iter_num=0)
    np.random.seed(7)
    L_orc, perms, _ = cluster_assignments_oracle(lv, la, N, 1, ind_groups, match, kdists=kd)
    assert tuple(L_ref.shape) == (N, hc)
    assert np.array_equal(L_ref.numpy(), L_orc), int((L_ref.numpy() != L_orc).sum())
    for h, head in enumerate(heads_a):
        w = list(head.modules())[-1].weight.data
        if match:
            assert torch.equal(w, w0[h][torch.from_numpy(perms[h])]), h      # the reference permuted the audio head likewise
        else:
            assert torch.equal(w, w0[h])
    assert m.training and m.return_features is False


class _Rec:
    def __init__(self):
        self.lines = []

    def info(self, msg, **_k):
        self.lines.append(str(msg))


def test_cluster_logging_matches_reference(gloo_group):
    """`cluster()` (src/sk_utils.py:23-134): the NMI / adjusted-NMI / entropy / purity log lines of the mirror
    (selavi_b200.sk_utils._log_label_metrics) equal the reference's for the same old and new labels (10th call: the
    entropy / purity block is active)."""
    from selavi_b200.sk_utils import _log_label_metrics
    ref = ref_loader.load_sk_module("cpu", sweep=True)
    hc, K, N = 2, 8, 96
    ds = _Clips(N)
    ds._labels = np.random.default_rng(3).integers(0, 5, N).tolist()
    ds.valid_indices = np.arange(N)
    torch.manual_seed(31)
    m = OracleAVModel(hc, K)
    m.use_mlp = True
    m.train()
    with torch.no_grad():
        for it in range(10):
            m(ds.v[it * 8:it * 8 + 16], ds.a[it * 8:it * 8 + 16])
    args = types.SimpleNamespace(world_size=1, rank=0, workers=0, ind_groups=1, headcount=hc, match=False, distribution="default",
                                 dist=None, diff_dist_every=False, diff_dist_per_head=True, gauss_sd=0.1, lamb=20, dump_path="")
    old = torch.from_numpy(np.random.default_rng(4).integers(0, K, (N, hc)))
    log_ref = _Rec()
    np.random.seed(1)
    torch.manual_seed(1)
    new = ref.cluster(args, old.clone(), ds, _DDP(m), 9, log_ref, None, None, 3)
    log_mine = _Rec()
    _log_label_metrics(args, new, old, ds, 10, log_mine, None, 3)
    keys = ("NMI_v:", "NMI-tolabels:", "aNMI-tolabels:", "Avg entropy:", "Avg purity:")
    pick = lambda lines: [l for l in lines if l.startswith(keys)]      # noqa: E731
    assert len(pick(log_ref.lines)) == 5
    assert pick(log_mine.lines) == pick(log_ref.lines)
