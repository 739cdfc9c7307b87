"""Pin down tcgen05.mma.kind::tf32 operand-layout conventions on real hardware (run on the GPU box).

For each hypothesis the host builds the raw shared-memory images of A [128 x K] and B [N x K] (K = 8 per MMA),
the descriptor bits and per-instruction offsets, runs selavi_debug_umma_probe and compares with A @ B^T.
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from selavi_b200 import _lib  # noqa: E402
from tools.probe import probe_lib  # noqa: E402


def desc_bits(lbo, sbo, layout):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | (layout & 7) << 61


def idesc(M, N, amaj, bmaj):
    return (1 << 4) | (2 << 7) | (2 << 10) | (amaj << 15) | (bmaj << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def img_kmajor_sw128(X, nk):
    """X [R, 8*nk] -> K-major SW128: row r at r*128, 16B chunk c at (c ^ (r&7)); 32 floats per 128B row."""
    R, K = X.shape
    assert K <= 32
    img = np.zeros((R, 32), np.float32)
    for r in range(R):
        for c in range(8):
            src = X[r, c * 4:(c + 1) * 4] if (c + 1) * 4 <= K else np.zeros(4, np.float32)
            pc = c ^ (r & 7)
            img[r, pc * 4:pc * 4 + 4] = src
    offs = [k * 32 for k in range(nk)]
    return img.reshape(-1), offs, desc_bits(16, 1024, 2)


def img_mnmajor_sw128(X, nk, swap=False):
    """X [R(mn), K]: atom(kg, mchunk) at (kg*NCH + mchunk)*1024; row = k%8; chunk (mn%32)/4 ^ (k%8)."""
    R, K = X.shape
    nch = (R + 31) // 32
    img = np.zeros((K // 8, nch, 8, 32), np.float32)
    for k in range(K):
        for mn in range(R):
            c = (mn % 32) // 4
            pc = c ^ (k % 8)
            img[k // 8, mn // 32, k % 8, pc * 4 + mn % 4] = X[mn, k]
    offs = [kg * nch * 1024 for kg in range(nk)]
    lbo, sbo = 1024, nch * 1024
    if swap:
        lbo, sbo = sbo, lbo
    return img.reshape(-1), offs, desc_bits(lbo, sbo, 2)


def img_mnmajor_sw128_kinner(X, nk, swap=False):
    """same atoms, but atoms ordered [mchunk][kg] (M-chunk stride = nk*1024, k-group stride = 1024)."""
    R, K = X.shape
    nch = (R + 31) // 32
    img = np.zeros((nch, K // 8, 8, 32), np.float32)
    for k in range(K):
        for mn in range(R):
            pc = ((mn % 32) // 4) ^ (k % 8)
            img[mn // 32, k // 8, k % 8, pc * 4 + mn % 4] = X[mn, k]
    offs = [kg * 1024 for kg in range(nk)]
    lbo, sbo = (K // 8) * 1024, 1024
    if swap:
        lbo, sbo = sbo, lbo
    return img.reshape(-1), offs, desc_bits(lbo, sbo, 2)


def img_mnmajor_noswz(X, nk, swap=False):
    """no-swizzle MN-major: core matrix = 8 k x 16B (4 mn) contiguous 128B; cores contiguous along MN, k-groups after."""
    R, K = X.shape
    nc = R // 4
    img = np.zeros((K // 8, nc, 8, 4), np.float32)
    for k in range(K):
        for mn in range(R):
            img[k // 8, mn // 4, k % 8, mn % 4] = X[mn, k]
    offs = [kg * nc * 128 for kg in range(nk)]
    sbo, lbo = 128, nc * 128   # canonical ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO))
    if swap:
        lbo, sbo = sbo, lbo
    return img.reshape(-1), offs, desc_bits(lbo, sbo, 0)


def run(a_img, a_offs, a_bits, b_img, b_offs, b_bits, idesc_v, N, dev):
    lib = probe_lib()
    a = torch.from_numpy(np.ascontiguousarray(a_img)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(b_img)).to(dev)
    ao = torch.tensor(a_offs, dtype=torch.int32, device=dev)
    bo = torch.tensor(b_offs, dtype=torch.int32, device=dev)
    out = torch.zeros(128, N, dtype=torch.float32, device=dev)
    pad = lambda t: (t.numel() * 4 + 15) // 16 * 16
    code = lib.selavi_debug_umma_probe(_lib.ptr(a), pad(a), _lib.ptr(b), pad(b), ctypes.c_ulonglong(a_bits),
                                       ctypes.c_ulonglong(b_bits), idesc_v, len(a_offs), _lib.ptr(ao), _lib.ptr(bo), N, 0,
                                       _lib.ptr(out), _lib.stream_ptr())
    assert code == 0, f"probe failed with code {code}"
    torch.cuda.synchronize()
    return out.cpu().numpy()


def trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def main(only=None):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    N, nk = 64, 2
    K = 8 * nk
    A = trunc(rng.standard_normal((128, K)).astype(np.float32))
    B = trunc(rng.standard_normal((N, K)).astype(np.float32))
    ref = A.astype(np.float64) @ B.astype(np.float64).T

    counter = [0]

    def want():
        counter[0] += 1
        return only is None or counter[0] - 1 == only

    def report(name, out):
        err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        print(f"{name:40s} rel={err:.3e} |out|={np.linalg.norm(out):.3e} |ref|={np.linalg.norm(ref):.3e} "
              f"nonzero={np.count_nonzero(out)}", flush=True)

    hyps_a = {"K": lambda X: img_kmajor_sw128(X, nk),
              "MN_sw128": lambda X: img_mnmajor_sw128(X, nk),
              "MN_sw128_swap": lambda X: img_mnmajor_sw128(X, nk, True),
              "MN_sw128_kin": lambda X: img_mnmajor_sw128_kinner(X, nk),
              "MN_sw128_kin_swap": lambda X: img_mnmajor_sw128_kinner(X, nk, True),
              "MN_nosw": lambda X: img_mnmajor_noswz(X, nk),
              "MN_nosw_swap": lambda X: img_mnmajor_noswz(X, nk, True)}
    # 1) sanity: both K-major
    ia, oa, ba = hyps_a["K"](A)
    ib, ob, bb = hyps_a["K"](B)
    if want():
        report("A=K B=K", run(ia, oa, ba, ib, ob, bb, idesc(128, N, 0, 0), N, dev))
    # 2) A MN-major variants with B K-major, and B MN-major variants with A K-major
    for name, f in hyps_a.items():
        if name == "K":
            continue
        if want():
            ia2, oa2, ba2 = f(A)
            report(f"A={name} B=K", run(ia2, oa2, ba2, ib, ob, bb, idesc(128, N, 1, 0), N, dev))
        if want():
            ib2, ob2, bb2 = f(B)
            report(f"A=K B={name}", run(ia, oa, ba, ib2, ob2, bb2, idesc(128, N, 0, 1), N, dev))
    # 3) both MN-major (the wgrad configuration)
    for name in ("MN_sw128", "MN_sw128_kin", "MN_nosw"):
        if want():
            ia2, oa2, ba2 = hyps_a[name](A)
            ib2, ob2, bb2 = hyps_a[name](B)
            report(f"A={name} B={name}", run(ia2, oa2, ba2, ib2, ob2, bb2, idesc(128, N, 1, 1), N, dev))


    return counter[0]


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(int(sys.argv[1]))
    else:
        import subprocess
        for i in range(16):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=90)
                out = [l for l in (r.stdout + r.stderr).splitlines() if "rel=" in l or "rror" in l]
                print(f"[{i}] rc={r.returncode} " + (" | ".join(out[-2:]) if out else "(no output)"), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT", flush=True)
