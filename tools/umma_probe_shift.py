"""tcgen05 probes for the halo (tap-reuse) convolution kernels.  Run on the GPU box: python tools/umma_probe_shift.py

 1. Row-shifted A windows: the A tile is a linear array of 128-byte rows (K-major SWIZZLE_128B, swizzle phase
    taken from the absolute row index), and the descriptor start address is advanced by j rows (j*128 bytes, not
    a multiple of the 1024-byte swizzle atom).  Tested with the descriptor's base-offset field (bits 49..51) = 0
    and = j & 7.  Expected result: A[j:j+128] @ B^T.
 2. Mixed operand types in kind::f16 (A fp16 x B bf16, and fp16 x fp16).
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from selavi_b200 import _lib  # noqa: E402
from tools.probe import probe_lib  # noqa: E402


def desc_bits(lbo, sbo, layout, base_off=0):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | (base_off & 7) << 49 | (layout & 7) << 61


def idesc_f16(M, N, afmt, bfmt):
    return (1 << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)


def bits16(x, fmt):
    t = torch.from_numpy(np.ascontiguousarray(x))
    t = t.half() if fmt == 0 else t.bfloat16()
    return t.view(torch.int16).numpy().view(np.uint16)


def rounded(x, fmt):
    t = torch.from_numpy(np.ascontiguousarray(x))
    return (t.half() if fmt == 0 else t.bfloat16()).float().numpy()


def img_rows_sw128(X, fmt):
    """linear rows of 128 B (64 elements); 16-byte chunk c of row r stored at chunk c ^ (r & 7)"""
    R, K = X.shape
    img = np.zeros((R, 64), np.uint16)
    b = bits16(X, fmt)
    for r in range(R):
        for c in range(K // 8):
            pc = c ^ (r & 7)
            img[r, pc * 8:pc * 8 + 8] = b[r, c * 8:c * 8 + 8]
    return img.reshape(-1)


def run(a_img, a_offs, a_bits, b_img, b_offs, b_bits, idesc_v, N, dev):
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


def case_shift(j, use_base_off, dev):
    rng = np.random.default_rng(1)
    N, nk = 64, 2
    K = 16 * nk
    A = rounded(rng.standard_normal((256, K)).astype(np.float32), 1)
    B = rounded(rng.standard_normal((N, K)).astype(np.float32), 1)
    ref = A[j:j + 128].astype(np.float64) @ B.astype(np.float64).T
    a_img = img_rows_sw128(A, 1)
    b_img = img_rows_sw128(B, 1)
    offs_a = [j * 128 + k * 32 for k in range(nk)]
    offs_b = [k * 32 for k in range(nk)]
    out = run(a_img, offs_a, desc_bits(16, 1024, 2, (j & 7) if use_base_off else 0), b_img, offs_b, desc_bits(16, 1024, 2),
              idesc_f16(128, N, 1, 1), N, dev)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    # which rows does the hardware actually deliver? match each output row against all candidate source rows
    full = A.astype(np.float64) @ B.astype(np.float64).T
    src = [int(np.argmin(np.linalg.norm(full - out[r], axis=1))) for r in (0, 1, 7, 8, 9, 127)]
    print(f"shift j={j:2d} base_off={'j&7' if use_base_off else '0  '} rel={err:.3e} rows(0,1,7,8,9,127)->{src}", flush=True)


def case_mixed(afmt, bfmt, dev):
    rng = np.random.default_rng(2)
    N, nk = 64, 2
    K = 16 * nk
    A = rounded(rng.standard_normal((128, K)).astype(np.float32), afmt)
    B = rounded(rng.standard_normal((N, K)).astype(np.float32), bfmt)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = run(img_rows_sw128(A, afmt), [k * 32 for k in range(nk)], desc_bits(16, 1024, 2), img_rows_sw128(B, bfmt),
              [k * 32 for k in range(nk)], desc_bits(16, 1024, 2), idesc_f16(128, N, afmt, bfmt), N, dev)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    print(f"mixed afmt={afmt} bfmt={bfmt} (0=fp16,1=bf16) rel={err:.3e}", flush=True)


CASES = [("shift", j, b) for j in (0, 1, 2, 3, 5, 8, 9, 58, 59) for b in (False, True)] + \
        [("mixed", 0, 0), ("mixed", 0, 1), ("mixed", 1, 0), ("mixed", 1, 1)]


def main(i):
    dev = torch.device("cuda:0")
    kind, x, y = CASES[i]
    if kind == "shift":
        case_shift(x, y, dev)
    else:
        case_mixed(x, y, dev)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "shift":
        for i, c in enumerate(CASES):     # wrong layouts only give wrong numbers: safe in one process
            if c[0] == "shift":
                main(i)
    elif len(sys.argv) > 1 and sys.argv[1] == "mn":
        pass
    elif len(sys.argv) > 1:
        main(int(sys.argv[1]))
    else:
        import subprocess
        r = subprocess.run([sys.executable, __file__, "shift"], capture_output=True, text=True, timeout=300)
        print(r.stdout + r.stderr[-2000:], flush=True)
        for i in [k for k, c in enumerate(CASES) if c[0] == "mixed"]:
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=90)
                out = [l for l in (r.stdout + r.stderr).splitlines() if "rel=" in l or "rror" in l]
                print(f"[{i}] rc={r.returncode} " + (" | ".join(out[-2:]) if out else "(no output)"), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT", flush=True)


# ------------------------------------------------------------------------------------------------------------------
# 3. MN-major A operand built from two row-shifted windows of one linear [k (pixel)][64 mn (channel)] array: the
#    weight-gradient GEMM with tap reuse.  M = 128 = window(q_a) | window(q_b); LBO = (q_b - q_a) * 128 bytes.
def case_mn_pair(qa, qb, dev):
    rng = np.random.default_rng(3)
    N, nk = 64, 2
    K = 16 * nk
    X = rounded(rng.standard_normal((256, 64)).astype(np.float32), 1)      # [pixel row][channel]
    B = rounded(rng.standard_normal((N, K)).astype(np.float32), 1)
    A = np.concatenate([X[qa:qa + K].T, X[qb:qb + K].T], 0)                # [128 (window, channel)][K pixels]
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a_img = img_rows_sw128(X, 1)            # row r at r*128, 16B chunk c at c ^ (r & 7): the MN-major SW128 atom layout
    b_img = img_rows_sw128(B, 1)
    offs_a = [(qa + 16 * k) * 128 for k in range(nk)]
    offs_b = [k * 32 for k in range(nk)]
    lbo = (qb - qa) * 128
    idesc = (1 << 4) | (1 << 7) | (1 << 10) | (1 << 15) | (0 << 16) | ((N >> 3) << 17) | ((128 >> 4) << 24)
    out = run(a_img, offs_a, desc_bits(lbo, 1024, 2), b_img, offs_b, desc_bits(16, 1024, 2), idesc, N, dev)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    e0 = np.linalg.norm(out[:64] - ref[:64]) / np.linalg.norm(ref[:64])
    e1 = np.linalg.norm(out[64:] - ref[64:]) / np.linalg.norm(ref[64:])
    print(f"mn_pair qa={qa:3d} qb={qb:3d} lbo={lbo:6d} rel={err:.3e} (rows 0-63 {e0:.2e}, rows 64-127 {e1:.2e})", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mn":
    for qa, qb in ((0, 8), (0, 64), (0, 1), (3, 4), (3, 61), (5, 63), (17, 133), (9, 9)):
        case_mn_pair(qa, qb, torch.device("cuda:0"))
