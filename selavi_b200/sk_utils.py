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

__all__ = ["optimize_L_sk_gpu", "optimize_L_sk_multi", "optimize_L_sk_sharded", "sk_solve_raw", "SKWorkspace"]


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
        self.sum = SymmetricBuffer(2 * ks * 8, group)
        self.flag = SymmetricBuffer(256, group)

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
