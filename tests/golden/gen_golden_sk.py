"""Generate tests/golden/sk_cases.npz by running the UNMODIFIED reference solver
(/root/reference/src/sk_utils.py:359 optimize_L_sk_gpu, device strings substituted to CPU by
oracle/ref_loader.py).  Run in the authoring container:  python tests/golden/gen_golden_sk.py
Inputs are regenerated from seeds by oracle.sk_oracle.synth_PS, so only outputs are stored.
"""
import os
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.sk_oracle import synth_PS  # noqa: E402

# name: (N, K, scale, seed, distribution, per_head, headcount, hc, lamb)
CASES = {
    "default_n512_k28": (512, 28, 1.0, 1, "default", True, 1, 0, 20.0),
    "default_n1001_k309_odd": (1001, 309, 1.0, 2, "default", True, 1, 0, 20.0),
    "default_n3000_k309_peaky": (3000, 309, 4.0, 3, "default", True, 1, 0, 20.0),
    "gauss_perhead_n2048_k32": (2048, 32, 1.0, 4, "gauss", True, 3, 1, 20.0),
    "gauss_shared_n1500_k400": (1500, 400, 2.0, 5, "gauss", False, 1, 0, 20.0),
    "default_n64_k28_cfg1": (64, 28, 1.0, 6, "default", True, 1, 0, 20.0),
    "gauss_perhead_n4000_k309": (4000, 309, 1.0, 7, "gauss", True, 10, 9, 20.0),
    "default_n777_k13_lamb8": (777, 13, 1.0, 8, "default", True, 1, 0, 8.0),
}


class _Log:
    def __init__(self):
        self.lines = []

    def info(self, s):
        self.lines.append(str(s))


def kdist_for(case):
    N, K, scale, seed, distribution, per_head, headcount, hc, lamb = CASES[case]
    if distribution == "default":
        return None
    rng = np.random.default_rng(1000 + seed)
    if per_head:
        return [(rng.standard_normal((K, 1)) * 0.1 + 1) * N / K for _ in range(headcount)]
    return np.maximum((rng.standard_normal((K, 1)) * 0.1 + 1) * N / K, 1.0)


def main():
    sk = ref_loader.load_sk_module("cpu")
    out = {}
    for name, (N, K, scale, seed, distribution, per_head, headcount, hc, lamb) in CASES.items():
        PS = torch.from_numpy(synth_PS(N, K, scale, seed))
        kd = kdist_for(name)
        args = types.SimpleNamespace(distribution=distribution, diff_dist_every=False, diff_dist_per_head=per_head,
                                     gauss_sd=0.1, headcount=headcount, lamb=lamb, rank=0, dist=None)
        if kd is not None:
            args.dist = [torch.from_numpy(k.copy()) for k in kd] if per_head else torch.from_numpy(kd.copy())
        log = _Log()
        cost, L = sk.optimize_L_sk_gpu(args, PS, hc, logger=log)
        m = re.search(r"error: ([^,]+), step : (\d+)", "\n".join(log.lines))
        out[name + "/labels"] = L.numpy().astype(np.int16)
        out[name + "/cost"] = np.float64(cost)
        out[name + "/iters"] = np.int64(m.group(2))
        out[name + "/err"] = np.float64(m.group(1))
        if kd is not None:
            after = args.dist[hc] if per_head else args.dist
            out[name + "/kdist_after"] = after.numpy().reshape(-1)
        print(name, "iters", m.group(2), "err", m.group(1), "cost", cost, flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sk_cases.npz"), **out)


if __name__ == "__main__":
    main()
