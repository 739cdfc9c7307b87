"""Multi-GPU parity worker: one process per GPU under torchrun (NCCL), started by tests/test_multigpu.py

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_worker.py <oracle.npz>

1. row-sharded Sinkhorn-Knopp (in-kernel NVSwitch exchange of the column sums) == CPU oracle on the full matrix, and the
   NCCL gather fallback (`optimize_L_sk_gathered`) gives the same labels
2. DDP + SyncBN train step on a rank-sharded batch AND the single-GPU step on the full batch: against each other and BOTH against the float64 CPU oracle of the full batch (global max(5e-3, 4 x the fp32 reference's own
   error); per parameter max(5e-2, 8 x the reference's own error) — see the comment at the comparison)
3. row-sharded dataset sweep + label assignment (`get_cluster_assignments_gpu`, cfg-4 flow: match, ind_groups = 2) ==
   CPU bookkeeping oracle, identical on every rank
Prints `MGPU_CHECK PASS|FAIL` on rank 0; exit code 1 on failure.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.sk_oracle import cluster_assignments_oracle, optimize_L_sk, synth_PS  # noqa: E402
from selavi_b200 import model as sv_model  # noqa: E402
from selavi_b200.sk_utils import (SKComm, get_cluster_assignments_gpu, optimize_L_sk_gathered,  # noqa: E402
                                  optimize_L_sk_sharded)
from selavi_b200.utils import get_loss  # noqa: E402

HC, K = 3, 309


def ddp_inputs(n=8):
    """n independent clips with per-clip gain / offset (tests/golden/gen_golden_model.py explains why)"""
    r = np.random.default_rng(11)
    gains = np.linspace(0.5, 2.0, n, dtype=np.float32)
    video = r.standard_normal((n, 3, 4, 64, 64)).astype(np.float32) * gains.reshape(n, 1, 1, 1, 1)
    spec = (r.standard_normal((n, 1, 257, 40)) * 17.89 + 1.93).astype(np.float32) * gains.reshape(n, 1, 1, 1)
    labels = r.integers(0, K, (n, HC)).astype(np.int64)
    return video, spec, labels


def build_model(factory):
    torch.manual_seed(31)
    m = factory()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


class Clips(torch.utils.data.Dataset):
    def __init__(self, n):
        r = np.random.default_rng(5)
        self.v = torch.from_numpy(r.standard_normal((n, 3, 4, 32, 32)).astype(np.float32) * np.linspace(0.5, 2, n, dtype=np.float32).reshape(n, 1, 1, 1, 1))
        self.a = torch.from_numpy((r.standard_normal((n, 1, 65, 40)) * 17.89 + 1.93).astype(np.float32))

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i], self.a[i], 0, i, i


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))   # a hang must fail fast
    gold = np.load(sys.argv[1])
    ok = True

    def say(msg):
        print(f"[rank {rank}/{world}] {msg}", flush=True)

    # ---- 1. sharded SK (cfg-5 shape at reduced N, and the K = 28 single-column-group instantiation with marginals)
    for (N, Ksk, scale, dist_kind) in [(4000, 309, 1.0, "default"), (6001 // world * world, 28, 2.0, "gauss"), (3328, 400, 1.0, "gauss")]:
        PS = synth_PS(N, Ksk, scale, seed=N)
        kd = (np.random.default_rng(1).standard_normal(Ksk) * 0.1 + 1) * N / Ksk if dist_kind == "gauss" else None
        ora = optimize_L_sk(PS, kdist=kd)
        n_local = N // world
        shard = torch.from_numpy(PS[rank * n_local:(rank + 1) * n_local]).to(dev)

        def mk_args():
            return types.SimpleNamespace(distribution=dist_kind, diff_dist_every=False, diff_dist_per_head=False, gauss_sd=0.1, headcount=1,
                                         lamb=20.0, rank=rank, dist=None if kd is None else torch.from_numpy(kd.copy()).view(Ksk, 1).to(dev))
        comm = SKComm(Ksk)
        cost, labels = optimize_L_sk_sharded(mk_args(), shard.clone(), 0, N, comm)
        same = np.array_equal(labels.cpu().numpy(), ora["labels"][rank * n_local:(rank + 1) * n_local])
        cost_ok = abs(cost - ora["cost"]) <= 1e-9 * abs(ora["cost"])
        cost_g, labels_g = optimize_L_sk_gathered(mk_args(), shard.clone(), 0)
        same_g = np.array_equal(labels_g.cpu().numpy(), ora["labels"][:n_local * world]) and abs(cost_g - ora["cost"]) <= 1e-9 * abs(ora["cost"])
        say(f"sharded SK N={N} K={Ksk} {dist_kind}: labels_equal={same} cost_ok={cost_ok} gathered_fallback_equal={same_g}")
        ok &= same and cost_ok and same_g

    # ---- 2. DDP + SyncBN and single GPU, both against the float64 oracle of the full batch
    video, spec, labels = ddp_inputs()
    B = video.shape[0]
    assert B % world == 0

    def factory():
        return sv_model.load_model(use_mlp=True, headcount=HC, num_classes=K, norm_feat=False)

    def step(net, v, s, lab):
        fv, fa = net(v, s)
        loss = 0.5 * get_loss(fv, lab, HC) + 0.5 * get_loss(fa, lab, HC)
        net.zero_grad()
        loss.backward()
        return loss.detach()

    single = build_model(factory).to(dev).train()
    l_single = step(single, torch.from_numpy(video).to(dev), torch.from_numpy(spec).to(dev), torch.from_numpy(labels).to(dev))
    g_single = {n: p.grad.clone() for n, p in single.named_parameters()}
    ddp_m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(build_model(factory)).to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(ddp_m, device_ids=[local], find_unused_parameters=True)   # main.py:156-160
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    l_ddp = step(ddp, torch.from_numpy(video[sl]).to(dev), torch.from_numpy(spec[sl]).to(dev), torch.from_numpy(labels[sl]).to(dev))
    dist.all_reduce(l_ddp)
    l_ddp /= world
    # the same step with DistributedDataParallel's stock gradient path (engine.DDP_BYPASS off: DDP reduces every parameter
    # and broadcasts the buffers itself): identical forward, so the gradients must agree to all-reduce rounding
    from selavi_b200 import engine
    engine.DDP_BYPASS = False
    stock_m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(build_model(factory)).to(dev).train()
    assert not hasattr(stock_m, "_ddp_params_and_buffers_to_ignore")
    stock = torch.nn.parallel.DistributedDataParallel(stock_m, device_ids=[local], find_unused_parameters=True)
    step(stock, torch.from_numpy(video[sl]).to(dev), torch.from_numpy(spec[sl]).to(dev), torch.from_numpy(labels[sl]).to(dev))
    engine.DDP_BYPASS = True
    g_ddp = dict(ddp_m.named_parameters())
    worst_stock = max((float((p.grad - g_ddp[n].grad).norm() / (p.grad.norm() + 1e-30)), n) for n, p in stock_m.named_parameters())
    say(f"engine-side gradient averaging vs stock DDP reducer: worst parameter {worst_stock[0]:.2e} ({worst_stock[1]}); "
        f"towers ignored by DDP: {len(ddp.parameters_to_ignore)} names")
    ok &= worst_stock[0] < 1e-5 and len(ddp.parameters_to_ignore) > 300 and len(stock.parameters_to_ignore) == 0
    # Three comparisons, per parameter and over the concatenated gradient ("global"):
    #   (a) DDP + SyncBN on the rank-sharded batch  vs  the same kernels on one GPU with the full batch: only the summation
    #       order of the BatchNorm statistics differs (1e-7 relative), but on this ill-conditioned toy step even that can
    #       flip a ReLU at the top of the towers: measured 1.7e-5 (no flip) or 1.2e-2 (one flip) depending on the kernel
    #       version, so the bar is the same 5e-2 as against float64.  The EXACT check of the multi-GPU plumbing is the
    #       bit-identity of the engine-side gradient averaging with DDP's stock reducer above;
    #   (b), (c) each of the two vs the float64 CPU oracle.  The clips are tiny (8 x 3x4x64x64: layer 4 holds 32 values per
    #       channel, BatchNorm1d sees 8 rows), so ONE ReLU whose pre-activation sits within the forward's 1e-4 rounding of
    #       zero changes a BatchNorm-bias gradient by percents (the fp32 reference, 10x more accurate in the forward, flips
    #       none).  The step is ill-conditioned for everyone: the fp32 REFERENCE is 9.2e-3 off float64 over the concatenated
    #       gradient (1 % on the early conv weights); the kernels here measure 1.9e-2, i.e. 2.1x the reference.  Bars: global
    #       max(5e-3, 4 x the reference's own global error), per parameter max(5e-2, 8 x the reference's own error).  The
    #       printed numbers are the evidence; round 1's "failure at >= 4 ranks" was this conditioning, not the exchange.
    rows = []
    num = {"ddp": 0.0, "single": 0.0, "pair": 0.0}
    den = 0.0
    worst_pair = (0.0, "")
    for n, p in ddp_m.named_parameters():
        g64 = torch.from_numpy(gold["grad64/" + n]).to(dev)
        gd, gs = p.grad.double(), g_single[n].double()
        d2 = float(g64.pow(2).sum())
        den += d2
        num["ddp"] += float((gd - g64).pow(2).sum())
        num["single"] += float((gs - g64).pow(2).sum())
        num["pair"] += float((gd - gs).pow(2).sum())
        e_ddp = float((gd - g64).norm()) / (d2 ** 0.5 + 1e-30)
        e_single = float((gs - g64).norm()) / (d2 ** 0.5 + 1e-30)
        e_pair = float((gd - gs).norm()) / (float(gs.norm()) + 1e-30)
        if e_pair > worst_pair[0]:
            worst_pair = (e_pair, n)
        bar = max(5e-2, 8 * float(gold["err32/" + n]))
        rows.append((max(e_ddp, e_single) / bar, n, e_ddp, e_single, float(gold["err32/" + n]), bar))
    rows.sort(reverse=True)
    glob = {k: (v / den) ** 0.5 for k, v in num.items()}
    loss64 = float(gold["loss64"])
    loss_ok = abs(float(l_ddp) - loss64) < 1e-4 * abs(loss64) and abs(float(l_single) - loss64) < 1e-4 * abs(loss64)
    gbar = max(5e-3, 4 * float(gold["global_err32"]))
    grads_ok = rows[0][0] <= 1.0 and glob["ddp"] < gbar and glob["single"] < gbar and worst_pair[0] < 5e-2
    say(f"DDP+SyncBN: loss ddp {float(l_ddp):.6f} single {float(l_single):.6f} fp64 oracle {loss64:.6f}; global gradient error vs fp64: "
        f"ddp {glob['ddp']:.2e} single {glob['single']:.2e} (reference fp32 {float(gold['global_err32']):.2e}, bar {gbar:.2e}); ddp vs single: global {glob['pair']:.2e}, worst parameter "
        f"{worst_pair[0]:.2e} ({worst_pair[1]})")
    if rank == 0:
        print("    worst parameters vs fp64 (relative to their bar):", flush=True)
        for _, n, ed, es, e32, bar in rows[:6]:
            print(f"    {n}: ddp {ed:.2e} single {es:.2e} ref32 {e32:.2e} bar {bar:.2e}", flush=True)
        bnb = [(ed, es, n) for _, n, ed, es, _, _ in rows if n.endswith(".bias") and ("bn" in n or ".1.bias" in n or ".4.bias" in n)]
        print(f"    BatchNorm-bias gradients: worst ddp {max(b[0] for b in bnb):.2e}, worst single {max(b[1] for b in bnb):.2e} "
              f"over {len(bnb)} tensors", flush=True)
    # running statistics: SyncBN uses the global batch => identical to the single-GPU full-batch model
    sd_s, sd_d = single.state_dict(), ddp_m.state_dict()
    buf_err = max(float((sd_s[k].float() - sd_d[k].float()).norm() / (sd_s[k].float().norm() + 1e-12)) for k in sd_s if "running" in k)
    say(f"running statistics ddp vs single: worst rel err {buf_err:.2e}")
    ok &= loss_ok and grads_ok and buf_err < 1e-5

    # ---- 3. row-sharded sweep + assignment, cfg-4 flow (match, ind_groups=2, K=28, N % world != 0)
    hc4, K4, N4 = 4, 28, 203
    ds = Clips(N4)
    m1 = build_model(lambda: sv_model.load_model(use_mlp=True, headcount=hc4, num_classes=K4, norm_feat=False)).to(dev)
    m1.train()
    with torch.no_grad():      # BN warm-up (utils.py:389-418): without it ~40 % of match_order's swap deltas are exact ties
        for it in range(25):   # (see tests/test_sweep.py:_warm_bn); same data on every rank => identical models
            lo = (it * 32) % 160
            m1(ds.v[lo:lo + 40].to(dev), ds.a[lo:lo + 40].to(dev))
    m1.eval()
    with torch.no_grad():
        m1.return_features = True
        fv, fa = m1(ds.v.to(dev), ds.a.to(dev))
        m1.return_features = False
        lv = [getattr(m1, f"mlp_v{h}").forward(fv).cpu().numpy() for h in range(hc4)]
        la = [getattr(m1, f"mlp_a{h}").forward(fa).cpu().numpy() for h in range(hc4)]
    m1.train()
    kd = [(np.random.default_rng(20 + h).standard_normal(K4) * 0.1 + 1) * N4 / K4 for h in range(hc4)]
    m2 = torch.nn.parallel.DistributedDataParallel(m1, device_ids=[local], find_unused_parameters=True)
    margs = types.SimpleNamespace(world_size=world, rank=rank, workers=0, ind_groups=2, headcount=hc4, match=True, distribution="gauss",
                                  dist=[torch.from_numpy(k.copy()).view(K4, 1).to(dev) for k in kd], diff_dist_every=False,
                                  diff_dist_per_head=True, gauss_sd=0.1, lamb=20.0, dump_path="")
    np.random.seed(0)
    L2 = get_cluster_assignments_gpu(margs, ds, m2, logger=None, iter_num=0)
    np.random.seed(0)
    L_ref, perms, _ = cluster_assignments_oracle(lv, la, N4, world, 2, True, kdists=kd)
    same = np.array_equal(L2.cpu().numpy(), L_ref)
    Lsum = L2.sum().clone()
    dist.all_reduce(Lsum)
    same_everywhere = int(Lsum) == world * int(L2.sum())
    say(f"sharded sweep (match, ind_groups=2, N={N4}, {N4 % world} remainder rows): labels == oracle {same}, identical on all ranks {same_everywhere}")
    ok &= same and same_everywhere

    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
