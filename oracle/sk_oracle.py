"""CPU oracle (numpy, float64) of the reference Sinkhorn-Knopp solver.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/sk_utils.py:359-422 (`optimize_L_sk_gpu`) line by line; the random draw of
the Gaussian marginals (sk_utils.py:372,377) is an INPUT here (`kdist`) so that the oracle, the reference
and the CUDA kernel can share it.  Pinned against the reference itself (device strings substituted, see
oracle/ref_loader.py) by tests/golden/gen_golden.py -> tests/golden/sk_*.npz, checked in
tests/test_oracle.py.  `cluster_assignments_oracle` / `match_order_oracle` (the sweep bookkeeping and the head
alignment, src/sk_utils.py:137-467) are pinned by tests/test_reference_bookkeeping.py, which runs the unmodified
reference functions on the CPU next to them.
"""
import numpy as np


def sk_marginals(PS, kdist):
    """sk_utils.py:369,388-392: scatter the target sizes by the argsort of the current cluster masses.

    NOTE the reference quirk: `_K_dist` is [K,1] and `torch.sort(_K_dist)` sorts along the LAST dimension
    (size 1), i.e. it is the identity; line 388 therefore computes  new[argsort[i]] = old[i]  (a permutation
    scatter, no sorting).  Verified against the reference itself (tests/golden/sk_cases.npz, kdist_after).

    PS [N,K] f64 (un-powered), kdist [K] f64 or None ('default' distribution: ones).
    Returns (r [K], kdist_permuted [K])."""
    K = PS.shape[1]
    if kdist is None:
        kd = np.ones(K, dtype=np.float64)
    else:
        kd = np.array(kdist, dtype=np.float64).reshape(K).copy()
        order = np.argsort(PS.sum(0), kind="stable")      # marginals_argsort (:369)
        kd[order] = kd.copy()                             # _K_dist[marginals_argsort] = sort(_K_dist, dim=-1)[0] (:388)
    r = 1.0 / kd                                          # (:392)
    r = r / r.sum()                                       # (:393)
    return r, kd


def optimize_L_sk(PS, lamb=20.0, kdist=None, max_iters=2000, check_every=10, tol=1e-1,
                  stop_on_converge=True):
    """Returns dict(cost, labels, alpha, beta, iters, err, kdist).  PS is not modified."""
    PS = np.array(PS, dtype=np.float64)
    N, K = PS.shape
    r, kd = sk_marginals(PS, kdist)
    beta = np.full(N, 1.0 / N)                            # (:390)
    PS = PS ** (0.5 * lamb)                               # (:391)
    c = 1.0 / N                                           # (:395)
    err = 1e6
    it = 0
    alpha = None
    while ((err > tol) if stop_on_converge else True) and it < max_iters:   # (:400)
        alpha = r / (beta @ PS)                           # (:401)
        beta_new = c / (PS @ alpha)                       # (:402)
        if it % check_every == 0:                         # (:403)
            err = float(np.sum(np.abs(beta / beta_new - 1.0)))
        beta = beta_new
        it += 1
    P = (PS * beta[:, None]) * alpha[None, :]             # (:411-412)
    labels = np.argmax(P, axis=1).astype(np.int64)        # (:413)
    P = ((1.0 / alpha)[None, :] * P) * (1.0 / beta)[:, None]   # (:416-417)
    with np.errstate(divide="ignore", invalid="ignore"):
        sol = np.nansum(np.log(P[np.arange(N), labels]))  # (:418)
    cost = -(1.0 / lamb) * sol / N                        # (:419)
    return dict(cost=float(cost), labels=labels, alpha=alpha, beta=beta, iters=it, err=err, kdist=kd)


def top2_margin(PS, lamb, alpha, beta):
    """Relative gap between the best and second-best entry of each row of the scaled matrix — used by the
    parity tests to tell a genuine mismatch from an ulp-level tie."""
    P = (np.array(PS, dtype=np.float64) ** (0.5 * lamb)) * beta[:, None] * alpha[None, :]
    part = np.partition(P, -2, axis=1)
    best, second = part[:, -1], part[:, -2]
    return (best - second) / np.maximum(best, 1e-300)


def synth_PS(N, K, scale=1.0, seed=0):
    """cfg-5 style input: softmax(randn*s) * softmax(randn*s), float64, numpy RNG."""
    rng = np.random.default_rng(seed)

    def sm(x):
        x = x - x.max(1, keepdims=True)
        e = np.exp(x)
        return e / e.sum(1, keepdims=True)

    return sm(rng.standard_normal((N, K)) * scale) * sm(rng.standard_normal((N, K)) * scale)


def match_order_oracle(emb1, emb2_in, steps=50000, restarts=2):
    """CPU restatement of the permutation search of src/sk_utils.py:424-467 (`match_order`): random pair swaps
    (np.random.choice stream) that reduce sum |emb1 - emb2[:, perm]|, early stop after 1000 non-improving steps,
    best of `restarts` tries if it beats the identity.  Returns the permutation applied to the audio head rows."""
    emb1 = np.asarray(emb1, dtype=np.float64)
    emb2_in = np.asarray(emb2_in, dtype=np.float64)
    K = emb1.shape[1]

    def c(a, b):
        return np.abs(a - b).sum()

    fin_perm = np.arange(K)
    last_iter = 0
    cost = c(emb1, emb2_in)
    best_cost = cost
    for _ in range(restarts):
        perm = np.arange(K)
        emb2 = emb2_in.copy()
        for _iter in range(steps):
            i, j = np.random.choice(K, 2, replace=False)
            current = c(emb1[:, i], emb2[:, i]) + c(emb1[:, j], emb2[:, j])
            future = c(emb1[:, i], emb2[:, j]) + c(emb1[:, j], emb2[:, i])
            if current - future > 0:
                emb2[:, [i, j]] = emb2[:, [j, i]]
                perm[i], perm[j] = perm[j], perm[i]
                last_iter = _iter
            if _iter - last_iter > 1000:
                break
        cost_try = c(emb1, emb2_in[:, perm])
        if cost_try < best_cost:
            best_cost = cost_try
            fin_perm = perm.copy()
    return fin_perm


def softmax64(x):
    """torch.nn.functional.softmax(x, dim=1, dtype=torch.float64) (src/sk_utils.py:206-211,272-275)."""
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - x.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


def cluster_assignments_oracle(logits_v, logits_a, N, world_size, ind_groups, match_first_iter, kdists=None, lamb=20.0):
    """CPU restatement of the BOOKKEEPING of `get_cluster_assignments_gpu` (src/sk_utils.py:183-327) for inputs that do
    not depend on the sweep order: `logits_v[h]`, `logits_a[h]` are the outputs of head h ([N, K], dataset-index order;
    for headcount 1 the model's own outputs).  Follows the reference's control flow and np.random consumption:

      order_heads = shuffle(range(hc))                                            (:183-184)
      for g in range(ind_groups):                                                 (:186)
          [sweep]                                                                 (:188-254)
          if match and iter_num == 0: for head in order_heads[g::ind_groups]:     (:257-286)
              perm = match_order(softmax64(v), softmax64(a)); permute the audio head's last Linear rows
              (= the COLUMNS of its output from now on)                           (:424-467)
          for head in order_heads[g::ind_groups]:                                 (:300-323)
              PS = softmax64(v) * softmax64(a); cost, L_head = SK(PS); L[indices, head] = L_head

    Rows `>= (N // world_size) * world_size` are never visited and keep label 0 (:157-161).
    Returns (L [N, hc] int64, perms {head: permutation}, order_heads)."""
    hc = len(logits_v)
    order_heads = list(range(hc))
    np.random.shuffle(order_heads)
    visited = (N // world_size) * world_size
    L = np.zeros((N, hc), dtype=np.int64)
    perms = {}
    la = [np.asarray(a, dtype=np.float32).copy() for a in logits_a]
    for g in range(ind_groups):
        heads = order_heads[g::ind_groups]
        if match_first_iter:
            for head in heads:
                perm = match_order_oracle(softmax64(logits_v[head][:visited]), softmax64(la[head][:visited]))
                if hc > 1:          # multi-head: head_a.forward(PS_a) is re-evaluated after the permutation (:309-312);
                    la[head] = la[head][:, perm]   # headcount 1 keeps using the outputs gathered BEFORE it (:301-303)
                perms[head] = perm
        for head in heads:
            PS = softmax64(logits_v[head][:visited]) * softmax64(la[head][:visited])
            out = optimize_L_sk(PS, lamb=lamb, kdist=None if kdists is None else kdists[head])
            L[:visited, head] = out["labels"]
    return L, perms, order_heads
