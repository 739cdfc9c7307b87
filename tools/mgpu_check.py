"""Multi-GPU checks, run with torchrun (NCCL), one process per GPU:
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
1. row-sharded Sinkhorn-Knopp (in-kernel NVSwitch P2P exchange of the column sums) == CPU oracle on the full matrix
2. DDP + SyncBN train step on a rank-sharded batch == single-GPU train step on the full batch (loss, gradients)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gen_golden_model import make_inputs  # noqa: E402
from oracle.sk_oracle import optimize_L_sk, synth_PS  # noqa: E402
from selavi_b200 import model as sv_model  # noqa: E402
from selavi_b200.sk_utils import SKComm, optimize_L_sk_sharded  # noqa: E402
from selavi_b200.utils import get_loss  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    # ---- 1. sharded SK
    for (N, K, scale, dist_kind) in [(4000, 309, 1.0, "default"), (6001 // world * world, 28, 2.0, "gauss")]:
        PS = synth_PS(N, K, scale, seed=N)
        kd = None
        if dist_kind == "gauss":
            kd = (np.random.default_rng(1).standard_normal(K) * 0.1 + 1) * N / K
        ora = optimize_L_sk(PS, kdist=kd)
        n_local = N // world
        shard = torch.from_numpy(PS[rank * n_local:(rank + 1) * n_local]).to(dev)
        args = types.SimpleNamespace(distribution=dist_kind, diff_dist_every=False, diff_dist_per_head=False, gauss_sd=0.1,
                                     headcount=1, lamb=20.0, rank=rank, dist=None if kd is None else torch.from_numpy(kd.copy()).view(K, 1).to(dev))
        comm = SKComm(K)
        cost, labels = optimize_L_sk_sharded(args, shard, 0, N, comm)
        same = np.array_equal(labels.cpu().numpy(), ora["labels"][rank * n_local:(rank + 1) * n_local])
        cost_ok = abs(cost - ora["cost"]) <= 1e-9 * abs(ora["cost"])
        print(f"[rank {rank}] sharded SK N={N} K={K} {dist_kind}: labels_equal={same} cost_ok={cost_ok} ({cost} vs {ora['cost']})", flush=True)
        ok &= same and cost_ok

    # ---- 2. DDP + SyncBN == single GPU on the full batch
    name = "mini_cfg2"
    video, spec, labels = make_inputs(name)
    if video.shape[0] % world or video.shape[0] < world:
        # more ranks than golden clips: independent synthetic clips with per-clip gains (scaled COPIES of the golden clips
        # make the heads' BatchNorm1d degenerate: r01r at 8 ranks, loss equal to 7 digits but 6.8e-3 on the worst gradient)
        r = np.random.default_rng(11)
        n = world * max(1, 4 // world)
        gains = np.linspace(0.5, 2.0, n, dtype=np.float32)
        video = (r.standard_normal((n,) + video.shape[1:]).astype(np.float32) * gains.reshape(n, 1, 1, 1, 1))
        spec = ((r.standard_normal((n,) + spec.shape[1:]) * 17.89 + 1.93).astype(np.float32) * gains.reshape(n, 1, 1, 1))
        labels = r.integers(0, 309, (n,) + labels.shape[1:]).astype(labels.dtype)
    if os.environ.get("MGPU_CLIPS"):   # experiment: fewer clips (e.g. one per rank)
        k = int(os.environ["MGPU_CLIPS"])
        video, spec, labels = video[:k], spec[:k], labels[:k]
    B = video.shape[0]
    hc, K = 3, 309

    def build():
        torch.manual_seed(31)
        m = sv_model.load_model(use_mlp=True, headcount=hc, num_classes=K, norm_feat=False)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return m

    def step(net, v, s, lab):
        fv, fa = net(v, s)
        loss = 0.5 * get_loss(fv, lab, hc) + 0.5 * get_loss(fa, lab, hc)
        net.zero_grad()
        loss.backward()
        return loss.detach()

    single = build().to(dev).train()
    l_single = step(single, torch.from_numpy(video).to(dev), torch.from_numpy(spec).to(dev), torch.from_numpy(labels).to(dev))
    g_single = {n: p.grad.clone() for n, p in single.named_parameters()}
    ddp_m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(build()).to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(ddp_m, device_ids=[local], find_unused_parameters=True)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    l_ddp = step(ddp, torch.from_numpy(video[sl]).to(dev), torch.from_numpy(spec[sl]).to(dev), torch.from_numpy(labels[sl]).to(dev))
    dist.all_reduce(l_ddp)
    l_ddp /= world
    worst, worst_name = 0.0, ""
    for n, p in ddp_m.named_parameters():
        e = float((p.grad - g_single[n]).norm() / (g_single[n].norm() + 1e-12))
        if e > worst:
            worst, worst_name = e, n
    print(f"[rank {rank}] DDP+SyncBN vs single GPU: loss {float(l_ddp):.6f} vs {float(l_single):.6f}, worst grad rel err {worst:.2e} ({worst_name})", flush=True)
    if rank == 0:
        errs = sorted(((float((p.grad - g_single[n]).norm() / (g_single[n].norm() + 1e-12)), n, float(g_single[n].norm()))
                       for n, p in ddp_m.named_parameters()), reverse=True)[:6]
        for e, n, gn in errs:
            print(f"    {e:.2e}  |g|={gn:.3e}  {n}", flush=True)
    ok &= abs(float(l_ddp) - float(l_single)) < 1e-4 * abs(float(l_single)) and worst < 5e-3
    # ---- 3. row-sharded dataset sweep + label assignment (get_cluster_assignments_gpu) == single-GPU result
    from selavi_b200.sk_utils import get_cluster_assignments_gpu

    class Clips(torch.utils.data.Dataset):
        def __init__(self, n):
            r = np.random.default_rng(5)
            self.v = torch.from_numpy(r.standard_normal((n, 3, 4, 32, 32)).astype(np.float32) * np.linspace(0.5, 2, n, dtype=np.float32).reshape(n, 1, 1, 1, 1))
            self.a = torch.from_numpy((r.standard_normal((n, 1, 65, 40)) * 17.89 + 1.93).astype(np.float32))

        def __len__(self):
            return len(self.v)

        def __getitem__(self, i):
            return self.v[i], self.a[i], 0, i, i

    ds = Clips(96)
    torch.manual_seed(31)
    m1 = sv_model.load_model(use_mlp=True, headcount=2, num_classes=8, norm_feat=False).to(dev)
    sargs = types.SimpleNamespace(world_size=1, rank=0, workers=0, ind_groups=1, headcount=2, match=False, distribution="default",
                                  dist=None, diff_dist_every=False, diff_dist_per_head=True, gauss_sd=0.1, lamb=20.0, dump_path="")
    import selavi_b200.sk_utils as sku
    np.random.seed(0)
    # single-process reference result (bypasses the process group on purpose)
    _init = dist.is_initialized
    dist.is_initialized = lambda: False
    L1 = get_cluster_assignments_gpu(sargs, ds, m1, logger=None)
    dist.is_initialized = _init
    m2 = torch.nn.parallel.DistributedDataParallel(m1, device_ids=[local], find_unused_parameters=True)
    margs = types.SimpleNamespace(**{**vars(sargs), "world_size": world, "rank": rank})
    np.random.seed(0)
    L2 = get_cluster_assignments_gpu(margs, ds, m2, logger=None)
    same = bool(torch.equal(L1, L2))
    print(f"[rank {rank}] sharded sweep labels == single-GPU labels: {same}", flush=True)
    ok &= same
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
