"""B200-native mirror of the reference `src/sk_utils.py` solver interface.

`optimize_L_sk_gpu(args, PS, hc, logger=None)` keeps the reference signature and side effects
(/root/reference/src/sk_utils.py:359-422): PS [N,K] float64 CUDA is consumed, `args.dist` is written /
permuted in place, `(cost: float, labels: LongTensor[N] cuda)` is returned.  The whole solve (pow, marginal
matching, iterations with the on-device stopping rule, argmax, cost) is ONE persistent CUDA kernel
(csrc/sk.cu) called through the C ABI `selavi_sk_solve`.  `optimize_L_sk_multi` is the north-star alias.

`optimize_L_sk_sharded` is the multi-GPU form: every rank keeps its row shard of PS, the only
cross-rank traffic is the K-vector of column sums read over NVSwitch P2P inside the kernel.

There is no CPU fallback: without the CUDA library these functions raise.
"""
import time

import torch

from . import _lib
from .symm import SymmetricBuffer

__all__ = ["optimize_L_sk_gpu", "optimize_L_sk_multi", "optimize_L_sk_sharded", "sk_solve_raw", "SKWorkspace", "SKComm",
           "optimize_L_sk_gathered", "softmax_product", "softmax64", "l1_cost_matrix", "get_cluster_assignments_gpu", "cluster", "match_order", "shard_range", "assemble_labels"]


class SKWorkspace:
    """Device scratch for selavi_sk_solve (per-CTA partial sums, marginals, barrier word, outputs)."""

    def __init__(self, K, n_local, device):
        lib = _lib.lib()
        nbytes = lib.selavi_sk_workspace_bytes(K)
        if nbytes == 0:
            raise _lib.SelaviError(f"Sinkhorn-Knopp kernel supports 1 <= K <= 512, got K={K}")
        self.K = K
        self.ws = torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device=device)
        self.alpha = torch.empty(K, dtype=torch.float64, device=device)
        self.beta = torch.empty(n_local, dtype=torch.float64, device=device)
        self.labels = torch.empty(n_local, dtype=torch.int64, device=device)
        self.iters = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.float64, device=device)
        self.cost = torch.zeros(1, dtype=torch.float64, device=device)


def sk_solve_raw(PS, n_global, lamb, kdist, ws, max_iters=2000, check_every=10, tol=1e-1, stop_on_converge=True,
                 do_prep=True, do_final=True, world=1, rank=0, peer_sum=None, peer_flag=None):
    """Thin checked wrapper over the C ABI; everything stays on the device (no sync)."""
    if not (PS.is_cuda and PS.dtype == torch.float64 and PS.dim() == 2 and PS.is_contiguous()):
        raise ValueError("PS must be a contiguous float64 CUDA matrix [N, K]")
    n_local, K = PS.shape
    if kdist is not None and not (kdist.is_cuda and kdist.dtype == torch.float64 and kdist.is_contiguous()
                                  and kdist.numel() == K):
        raise ValueError("kdist must be a contiguous float64 CUDA tensor with K elements")
    import ctypes
    ps_arr = pf_arr = None
    if world > 1:
        ps_arr = (ctypes.c_void_p * world)(*peer_sum)
        pf_arr = (ctypes.c_void_p * world)(*peer_flag)
    with torch.cuda.device(PS.device):
        code = _lib.lib().selavi_sk_solve(
            _lib.ptr(PS), n_local, n_global, K, float(lamb), 1 if kdist is not None else 0, _lib.ptr(kdist),
            _lib.ptr(ws.alpha), _lib.ptr(ws.beta), _lib.ptr(ws.labels), _lib.ptr(ws.ws), int(max_iters),
            int(check_every), float(tol), 1 if stop_on_converge else 0, 1 if do_prep else 0, 1 if do_final else 0,
            _lib.ptr(ws.iters), _lib.ptr(ws.err), _lib.ptr(ws.cost), world, rank, ps_arr, pf_arr, _lib.stream_ptr())
    _lib.check(code, "selavi_sk_solve")


def _pick_marginals(args, N, K, hc, device, logger):
    """Marginal bookkeeping of sk_utils.py:366-387 (host side; the draw uses torch's CUDA generator exactly
    like the reference so that seeds are interchangeable).  Returns the [K,1] tensor to be permuted in place,
    or None for the 'default' distribution."""
    if args.distribution == 'default':
        return None
    _K_dist = torch.ones((K, 1), dtype=torch.float64, device=device)
    if (args.dist is None) or args.diff_dist_every:
        if args.distribution == 'gauss':
            if args.diff_dist_per_head:
                _K_dists = [(torch.randn(size=(K, 1), dtype=torch.float64, device=device) * args.gauss_sd + 1) * N / K
                            for _ in range(args.headcount)]
                args.dist = _K_dists
                _K_dist = _K_dists[hc]
            else:
                _K_dist = (torch.randn(size=(K, 1), dtype=torch.float64, device=device) * args.gauss_sd + 1) * N / K
                _K_dist = torch.clamp(_K_dist, min=1)
                args.dist = _K_dist
        if args.rank == 0 and logger is not None:
            logger.info(f"distribution used: {_K_dist}")
    else:
        _K_dist = args.dist[hc] if args.diff_dist_per_head else args.dist
    if not _K_dist.is_contiguous():
        raise ValueError("args.dist entries must be contiguous")
    return _K_dist


def optimize_L_sk_gpu(args, PS, hc, logger=None):
    """Drop-in for src/sk_utils.py:359.  Returns (cost, newL)."""
    tt = time.time()
    N, K = PS.shape
    kdist = _pick_marginals(args, N, K, hc, PS.device, logger)
    ws = SKWorkspace(K, N, PS.device)
    sk_solve_raw(PS, N, args.lamb, kdist, ws)
    iters, err, sol = int(ws.iters.item()), float(ws.err.item()), float(ws.cost.item())  # one sync, like the reference
    cost = -(1. / args.lamb) * sol / N
    if args.rank == 0 and logger is not None:
        logger.info(f"error: {err}, step : {iters}")
        logger.info(f"opt took {(time.time() - tt) / 60.} min, {iters} iters")
    return cost, ws.labels


optimize_L_sk_multi = optimize_L_sk_gpu  # name used by BASELINE.json's north_star


class SKComm:
    """Symmetric P2P exchange buffers for the sharded solve (one per process group and K)."""

    def __init__(self, K, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        ks = _lib.lib().selavi_sk_kp(K)
        self.sum = SymmetricBuffer(2 * self.world * ks * 8, group)   # [2][world][Ks] receive slots
        self.flag = SymmetricBuffer(256, group)                       # [world] epoch flags

    def reset(self):
        """Zero the epoch flag on every rank before a solve (the kernel counts epochs from 0)."""
        import torch.distributed as dist
        self.flag.zero_()
        torch.cuda.synchronize()
        dist.barrier(self.group)


def optimize_L_sk_sharded(args, PS_local, hc, n_global, comm, logger=None, kdist=None):
    """Row-sharded solve: PS_local [N_r, K] f64 on this rank, sum_r N_r == n_global.

    Gaussian marginals are drawn on rank 0 and broadcast (the reference draws them on its single solver GPU).
    Returns (cost, labels_local); cost is the global cost (all-reduced)."""
    import torch.distributed as dist
    N_local, K = PS_local.shape
    if kdist is None:
        kdist = _pick_marginals(args, n_global, K, hc, PS_local.device, logger)
        if kdist is not None:
            dist.broadcast(kdist, dist.get_global_rank(comm.group, 0) if comm.group is not None else 0, group=comm.group)
    ws = SKWorkspace(K, N_local, PS_local.device)
    comm.reset()
    sk_solve_raw(PS_local, n_global, args.lamb, kdist, ws, world=comm.world, rank=comm.rank,
                 peer_sum=comm.sum.peer_ptrs, peer_flag=comm.flag.peer_ptrs)
    sol = ws.cost.clone()
    dist.all_reduce(sol, group=comm.group)
    cost = -(1. / args.lamb) * float(sol.item()) / n_global
    if args.rank == 0 and logger is not None:
        logger.info(f"error: {float(ws.err.item())}, step : {int(ws.iters.item())}")
    return cost, ws.labels


# ====================================================================================================
# Dataset-wide feature sweep + label assignment (src/sk_utils.py:23-356), row-sharded
# ====================================================================================================
def softmax_product(logits_v, logits_a):
    """float64 softmax(v) * softmax(a) in one pass (src/sk_utils.py:206-211,309-315)."""
    if not (logits_v.is_cuda and logits_v.dtype == torch.float32 and logits_a.shape == logits_v.shape):
        raise ValueError("softmax_product needs two float32 CUDA matrices of equal shape")
    logits_v, logits_a = logits_v.contiguous(), logits_a.contiguous()
    n, K = logits_v.shape
    PS = torch.empty((n, K), dtype=torch.float64, device=logits_v.device)
    with torch.cuda.device(PS.device):
        _lib.check(_lib.lib().selavi_sk_softmax_product(_lib.ptr(logits_v), _lib.ptr(logits_a), n, K, _lib.ptr(PS),
                                                        _lib.stream_ptr()), "selavi_sk_softmax_product")
    return PS


def shard_range(N, world_size, rank):
    """Rows owned by `rank` (src/sk_utils.py:157-161): N // world_size each, the remainder is dropped."""
    local = N // world_size
    return rank * local, (rank + 1) * local


def assemble_labels(L, idx_all, lab_all, head):
    """L[idx, head] = labels for the gathered (index, label) pairs of all ranks (src/sk_utils.py:323)."""
    L[idx_all.long(), head] = lab_all
    return L


_comm_cache = {}


def _sk_comm(K, group):
    """SKComm for (K, group), or None when the in-kernel NVSwitch exchange is unavailable (more than 8 ranks, ranks on
    several hosts, CUDA IPC refused).  The decision is collective — SymmetricBuffer raises on every rank or on none — so
    all ranks take the same branch."""
    import torch.distributed as dist
    key = (K, id(group))
    if key not in _comm_cache:
        comm = None
        if dist.get_world_size(group) <= 8:
            try:
                comm = SKComm(K, group)
            except _lib.SelaviError as e:
                import warnings
                warnings.warn(f"selavi_b200: P2P exchange for the sharded Sinkhorn-Knopp solve unavailable ({e}); gathering the "
                              "matrix to rank 0 over NCCL like the reference (src/sk_utils.py:214-242)")
        _comm_cache[key] = comm
    return _comm_cache[key]


def optimize_L_sk_gathered(args, PS_local, hc, group=None, logger=None):
    """Fallback of the sharded solve for any world size / multi-node (the reference's own data flow, src/sk_utils.py:
    214-242,287-327): the row shards are gathered on rank 0 over NCCL, rank 0 runs the single-GPU solver kernel and the
    labels are broadcast.  Returns (cost, labels of ALL rows in rank-major order) on every rank."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    root = dist.get_global_rank(group, 0) if group is not None else 0
    n_local, K = PS_local.shape
    shards = [torch.empty_like(PS_local) for _ in range(world)] if rank == 0 else None
    dist.gather(PS_local, shards, dst=root, group=group)
    labels = torch.empty(n_local * world, dtype=torch.int64, device=PS_local.device)
    cost = torch.zeros(1, dtype=torch.float64, device=PS_local.device)
    if rank == 0:
        c, lab = optimize_L_sk_gpu(args, torch.cat(shards), hc, logger=logger)
        labels.copy_(lab)
        cost[0] = c
    dist.broadcast(labels, root, group=group)
    dist.broadcast(cost, root, group=group)
    return float(cost.item()), labels


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def get_cluster_assignments_gpu(args, dataset, model, logger=None, writer=None, group=None, iter_num=0):
    """Same contract as src/sk_utils.py:137-356 (returns L [N, headcount] int64 on every rank, model left in train
    mode with return_features=False), different data flow: every rank keeps the features of ITS rows, applies the
    heads locally, and the ranks solve Sinkhorn-Knopp jointly on the row-sharded matrix — only the K-vector of
    column sums crosses NVSwitch (inside the solver kernel) and the labels are all-gathered at the end.  Nothing is
    gathered to rank 0, and there is no per-batch barrier."""
    import numpy as np
    import torch.distributed as dist
    from torch.utils.data.sampler import SubsetRandomSampler
    distributed = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if distributed else (1, 0)
    net = _unwrap(model)
    model.eval()
    N = len(dataset)
    lo, hi = shard_range(N, args.world_size, args.rank)
    sampler = SubsetRandomSampler(torch.arange(lo, hi).int())
    dataloader = torch.utils.data.DataLoader(dataset, batch_size=64, sampler=sampler, shuffle=False,
                                             num_workers=args.workers, pin_memory=True, collate_fn=None)
    if distributed:
        dist.barrier(group=group)
    assert args.ind_groups <= args.headcount
    hc = args.headcount
    if hc > 1:
        net.return_features = True
    dev = torch.device("cuda", torch.cuda.current_device())
    L = torch.zeros((N, hc), dtype=torch.long, device=dev)
    order_heads = list(range(hc))
    np.random.shuffle(order_heads)
    n_local = hi - lo
    for hd_grp_idx in range(args.ind_groups):
        # 1. sweep this rank's rows (fresh augmentations per head group, like the reference)
        feats_v, feats_a, idxs = [], [], []
        with torch.no_grad():
            for batch in dataloader:
                video, audio, _, idx, _ = batch
                fv, fa = model(video.cuda(non_blocking=True), audio.cuda(non_blocking=True))
                feats_v.append(fv.reshape(video.shape[0], -1))
                feats_a.append(fa.reshape(video.shape[0], -1))
                idxs.append(idx.cuda(non_blocking=True).long())
        F_v, F_a, idx_local = torch.cat(feats_v), torch.cat(feats_a), torch.cat(idxs)
        if args.match and iter_num == 0:
            for head in order_heads[hd_grp_idx::args.ind_groups]:
                head_a = net.mlp_a if hc == 1 else getattr(net, f"mlp_a{head}")
                head_v = net.mlp_v if hc == 1 else getattr(net, f"mlp_v{head}")
                with torch.no_grad():
                    Pv = F_v if hc == 1 else head_v.forward(F_v)
                    Pa = F_a if hc == 1 else head_a.forward(F_a)
                match_order(args, Pv, Pa, list(head_a.modules())[-1] if net.use_mlp else head_a, logger=logger, group=group,
                            logits=True)
        # 2. joint row-sharded Sinkhorn-Knopp per head
        _costs = [0 for _ in range(hc)]
        for head in order_heads[hd_grp_idx::args.ind_groups]:
            sk_start = time.time()
            with torch.no_grad():
                if hc == 1:
                    lv, la = F_v, F_a            # logits of the single head (the reference softmaxes them, :206-211)
                else:
                    lv = getattr(net, f"mlp_v{head}").forward(F_v)
                    la = getattr(net, f"mlp_a{head}").forward(F_a)
                PS = softmax_product(lv, la)
            if world > 1:
                comm = _sk_comm(PS.shape[1], group)
                gathered_idx = [torch.empty_like(idx_local) for _ in range(world)]
                dist.all_gather(gathered_idx, idx_local, group=group)
                if comm is not None:
                    cost, L_head = optimize_L_sk_sharded(args, PS, head, n_local * world, comm, logger=logger)
                    gathered_lab = [torch.empty_like(L_head) for _ in range(world)]
                    dist.all_gather(gathered_lab, L_head, group=group)
                    lab_all = torch.cat(gathered_lab)
                else:
                    cost, lab_all = optimize_L_sk_gathered(args, PS, head, group=group, logger=logger)
                assemble_labels(L, torch.cat(gathered_idx), lab_all, head)
            else:
                cost, L_head = optimize_L_sk_gpu(args, PS, hc=head, logger=logger)
                assemble_labels(L, idx_local, L_head, head)
            _costs[head] = cost
            if args.rank == 0 and logger is not None:
                logger.info(f"Head {head}, Cost: (video): {cost:.3f}; time: {time.time() - sk_start:.3f}")
        if args.rank == 0 and logger is not None:
            logger.info(f"Final Cost: (video): {np.mean(_costs):.3f}")
        if writer:
            writer.add_scalar('train/LP-cost', np.mean(_costs), iter_num)
    if distributed:
        dist.barrier(group=group)
    torch.cuda.synchronize()
    net.return_features = False
    model.train()
    return L


def softmax64(logits):
    """float64 softmax of fp32 head outputs (torch.nn.functional.softmax(x, dim=1, dtype=torch.float64),
    src/sk_utils.py:272-275) in one kernel."""
    if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2):
        raise ValueError("softmax64 needs a float32 CUDA matrix")
    logits = logits.contiguous()
    n, K = logits.shape
    out = torch.empty((n, K), dtype=torch.float64, device=logits.device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().selavi_sk_softmax64(_lib.ptr(logits), n, K, _lib.ptr(out), _lib.stream_ptr()), "selavi_sk_softmax64")
    return out


def l1_cost_matrix(P1, P2):
    """C[i, j] = sum_n |P1[n, i] - P2[n, j]|  (the quantity `c(a, b)` of src/sk_utils.py:430-431 for all column pairs),
    float64, one CUDA kernel + a fixed-order split-N reduce (csrc/match.cu)."""
    if not (P1.is_cuda and P1.dtype == torch.float64 and P2.dtype == torch.float64 and P1.shape == P2.shape and P1.dim() == 2):
        raise ValueError("l1_cost_matrix needs two float64 CUDA matrices of equal shape [N, K]")
    P1, P2 = P1.contiguous(), P2.contiguous()
    n, K = P1.shape
    lib = _lib.lib()
    C = torch.empty((K, K), dtype=torch.float64, device=P1.device)
    ws = torch.empty(lib.selavi_l1_cost_workspace_bytes(n, K) // 8, dtype=torch.float64, device=P1.device)
    with torch.cuda.device(P1.device):
        _lib.check(lib.selavi_l1_cost_matrix(_lib.ptr(P1), _lib.ptr(P2), n, K, _lib.ptr(C), _lib.ptr(ws), _lib.stream_ptr()),
                   "selavi_l1_cost_matrix")
    return C


@torch.no_grad()
def match_order(args, emb1, emb2_in, W2, steps=50000, restarts=2, logger=None, group=None, logits=False):
    """Same contract, RNG stream (np.random.choice) and result as src/sk_utils.py:424-467: random pair-swap
    hill-climb minimising sum |emb1 - emb2[:, perm]|, then the rows of the audio head's last Linear are permuted.
    Instead of ~10^5 x 15 tiny kernels with an .item() sync each, the K x K L1 cost matrix is computed once (rows
    sharded over ranks, all-reduced) and the hill-climb runs on it on the host.
    `logits=True`: inputs are this rank's head outputs; the float64 softmax (src/sk_utils.py:272-275) is applied here."""
    import numpy as np
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    if logits and emb1 is not None:
        emb1, emb2_in = softmax64(emb1), softmax64(emb2_in)
    K = len(W2.bias.data)
    C = l1_cost_matrix(emb1.double(), emb2_in.double())
    if distributed and dist.get_world_size(group) > 1:
        dist.all_reduce(C, group=group)
    fin_perm = np.arange(K)
    if args.rank == 0:
        Cn = C.cpu().numpy()
        ar = np.arange(K)
        cost = Cn[ar, ar].sum()
        best_cost = cost
        if logger is not None:
            logger.info(f'initial cost: {cost:.1f}')
        last_iter = 0
        for _ in range(restarts):
            perm = np.arange(K)
            for _iter in range(steps):
                i, j = np.random.choice(K, 2, replace=False)
                delta = (Cn[i, perm[i]] + Cn[j, perm[j]]) - (Cn[i, perm[j]] + Cn[j, perm[i]])
                if delta > 0:
                    perm[i], perm[j] = perm[j], perm[i]
                    last_iter = _iter
                if _iter - last_iter > 1000:
                    break
            cost_try = Cn[ar, perm].sum()
            if logger is not None:
                logger.info(f"cost of this try: {cost_try:.2f}")
            if cost_try < best_cost:
                best_cost = cost_try
                fin_perm = perm.copy()
        if logger is not None:
            logger.info(f"final cost: {best_cost:.2f}")
    fin = torch.from_numpy(fin_perm).to(W2.bias.device)
    if distributed and dist.get_world_size(group) > 1:
        dist.broadcast(fin, dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    W2.bias.data = W2.bias.data[fin]
    W2.weight.data = W2.weight.data[fin]
    return fin


def _log_label_metrics(args, new, old, dataset, sk_counter, logger, writer, iter_num):
    """Host-side bookkeeping around the assignment (src/sk_utils.py:36-118): NMI of the first head against the previous
    labels and against the dataset's ground truth, adjusted NMI, and every 10th call the per-cluster entropy / purity.
    Pure logging (scikit-learn / scipy on the CPU, as in the reference); skipped quietly when those are not installed."""
    import numpy as np
    try:
        from scipy.stats import entropy
        from sklearn.metrics.cluster import adjusted_mutual_info_score, normalized_mutual_info_score
    except ImportError:
        return
    info = logger.info if (logger is not None and args.rank == 0) else (lambda *_a, **_k: None)
    scalar = writer.add_scalar if writer else (lambda *_a, **_k: None)
    mine = new[:, 0].cpu().numpy()
    nmi_prev = normalized_mutual_info_score(mine, old[:, 0].cpu().numpy(), average_method='arithmetic')
    info(f'NMI_v: {nmi_prev}')
    scalar('train/nmi_v/iter', nmi_prev, iter_num)
    scalar('train/optim_count/iter', sk_counter, iter_num)
    truth = np.array(dataset._labels)[dataset.valid_indices]
    nmi_gt = normalized_mutual_info_score(mine, truth, average_method='arithmetic')
    anmi_gt = adjusted_mutual_info_score(mine, truth, average_method='arithmetic')
    info(f"NMI-tolabels: {nmi_gt}")
    info(f"aNMI-tolabels: {anmi_gt}")
    scalar('train/nmi-tolabels_v/iter', nmi_gt, iter_num)
    scalar('train/a-nmi-tolabels_v/iter', anmi_gt, iter_num)
    if sk_counter % 10 == 0:
        ents, purs = [], []
        for k in np.unique(mine):
            _, counts = np.unique(truth[mine == k], return_counts=True)
            frac = counts / counts.sum()
            purs.append(frac.max())
            ents.append(entropy(frac))
        if logger is not None:
            logger.info(f"Avg entropy: {np.mean(ents)}")
            logger.info(f"Avg purity: {np.mean(purs)}")
        if writer:
            writer.add_histogram('train/entropies', np.array(ents), iter_num)
            writer.add_histogram('train/purities', np.array(purs), iter_num)
            writer.add_scalar('train/avg-entropy', np.mean(ents), iter_num)
            writer.add_scalar('train/avg-purity', np.mean(purs), iter_num)


def cluster(args, selflabels, dataset, model, sk_counter, logger, writer, group, iter_num):
    """src/sk_utils.py:23-134: new pseudo-labels on every rank, the reference's NMI / purity logging (same logger lines
    and tensorboard tags), and the closing barrier.  The SLURM requeue hook (SIGNAL_RECEIVED, :120-125) is job control
    and stays with the reference's launcher."""
    import torch.distributed as dist
    old = selflabels.clone()
    with torch.no_grad():
        selflabels = get_cluster_assignments_gpu(args, dataset, model, logger, writer, group, iter_num)
    sk_counter += 1
    _log_label_metrics(args, selflabels, old, dataset, sk_counter, logger, writer, iter_num)
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)
    return selflabels
