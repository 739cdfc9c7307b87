"""Pin down tcgen05.mma.kind::f16 (bf16 operands) layout conventions: K-major and MN-major SWIZZLE_128B.
A [128 x K], B [N x K], K = 16 per MMA.  Run on the GPU box: python tools/umma_probe_bf16.py"""
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


def idesc_bf16(M, N, amaj, bmaj):
    return (1 << 4) | (1 << 7) | (1 << 10) | (amaj << 15) | (bmaj << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def to_bf16_bits(x):
    """float32 array (already bf16-representable) -> uint16 bit patterns"""
    return (x.view(np.uint32) >> 16).astype(np.uint16)


def img_k_sw128(X, nk):
    """K-major SW128, bf16: row r at r*128 B (64 elements), 16B chunk c (8 elements) at c ^ (r & 7)."""
    R, K = X.shape
    img = np.zeros((R, 64), np.uint16)
    bits = to_bf16_bits(X)
    for r in range(R):
        for c in range(K // 8):
            pc = c ^ (r & 7)
            img[r, pc * 8:pc * 8 + 8] = bits[r, c * 8:c * 8 + 8]
    return img.reshape(-1), [k * 32 for k in range(nk)], desc_bits(16, 1024, 2)


def img_mn_sw128(X, nk, swap=False, order="kg_outer"):
    """MN-major SW128, bf16: atom = 8 k-rows x 128 B (64 mn elements); chunk (mn%64)//8 ^ (k%8)."""
    R, K = X.shape
    nch = (R + 63) // 64
    bits = to_bf16_bits(X)
    if order == "kg_outer":
        img = np.zeros((K // 8, nch, 8, 64), np.uint16)
    else:
        img = np.zeros((nch, K // 8, 8, 64), np.uint16)
    for k in range(K):
        for mn in range(R):
            pc = ((mn % 64) // 8) ^ (k % 8)
            if order == "kg_outer":
                img[k // 8, mn // 64, k % 8, pc * 8 + mn % 8] = bits[mn, k]
            else:
                img[mn // 64, k // 8, k % 8, pc * 8 + mn % 8] = bits[mn, k]
    if order == "kg_outer":
        lbo, sbo = 1024, nch * 1024            # LBO: next 64-wide MN chunk, SBO: next 8-row K group
        offs = [i * 2 * nch * 1024 for i in range(nk)]
    else:
        lbo, sbo = (K // 8) * 1024, 1024
        offs = [i * 2 * 1024 for i in range(nk)]
    if swap:
        lbo, sbo = sbo, lbo
    return img.reshape(-1), offs, desc_bits(lbo, sbo, 2)


def run(a, b, idesc_v, N, dev):
    (a_img, a_offs, a_bits), (b_img, b_offs, b_bits) = a, b
    lib = probe_lib()
    at = torch.from_numpy(np.ascontiguousarray(a_img).view(np.int16)).to(dev)
    bt = torch.from_numpy(np.ascontiguousarray(b_img).view(np.int16)).to(dev)
    ao = torch.tensor(a_offs, dtype=torch.int32, device=dev)
    bo = torch.tensor(b_offs, dtype=torch.int32, device=dev)
    out = torch.zeros(128, N, dtype=torch.float32, device=dev)
    pad = lambda t: (t.numel() * 2 + 15) // 16 * 16  # noqa: E731
    code = lib.selavi_debug_umma_probe(_lib.ptr(at), pad(at), _lib.ptr(bt), pad(bt), ctypes.c_ulonglong(a_bits),
                                       ctypes.c_ulonglong(b_bits), idesc_v, len(a_offs), _lib.ptr(ao), _lib.ptr(bo), N, 1,
                                       _lib.ptr(out), _lib.stream_ptr())
    assert code == 0, f"probe failed with code {code}"
    torch.cuda.synchronize()
    return out.cpu().numpy()


def main(only=None):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    N, nk = 64, 2
    K = 16 * nk
    A = torch.from_numpy(rng.standard_normal((128, K)).astype(np.float32)).bfloat16().float().numpy()
    B = torch.from_numpy(rng.standard_normal((N, K)).astype(np.float32)).bfloat16().float().numpy()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    cases = [("A=K B=K", lambda: (img_k_sw128(A, nk), img_k_sw128(B, nk), 0, 0))]
    for order in ("kg_outer", "mn_outer"):
        for swap in (False, True):
            tag = f"MN[{order}{',swap' if swap else ''}]"
            cases.append((f"A={tag} B=K", lambda o=order, s=swap: (img_mn_sw128(A, nk, s, o), img_k_sw128(B, nk), 1, 0)))
            cases.append((f"A=K B={tag}", lambda o=order, s=swap: (img_k_sw128(A, nk), img_mn_sw128(B, nk, s, o), 0, 1)))
            cases.append((f"A={tag} B={tag}", lambda o=order, s=swap: (img_mn_sw128(A, nk, s, o), img_mn_sw128(B, nk, s, o), 1, 1)))
    for i, (name, f) in enumerate(cases):
        if only is not None and i != only:
            continue
        a, b, am, bm = f()
        out = run(a, b, idesc_bf16(128, N, am, bm), N, dev)
        err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        print(f"{name:44s} rel={err:.3e} |out|={np.linalg.norm(out):.3e} nonzero={np.count_nonzero(out)}", flush=True)
    return len(cases)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(int(sys.argv[1]))
    else:
        import subprocess
        for i in range(13):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=90)
                out = [l for l in (r.stdout + r.stderr).splitlines() if "rel=" in l or "rror" in l]
                print(f"[{i}] rc={r.returncode} " + (" | ".join(out[-2:]) if out else "(no output)"), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT", flush=True)
