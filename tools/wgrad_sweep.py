"""Weight-gradient split-K schedule sweep (run on the GPU box): SELAVI_WGRAD_WAVES x SELAVI_WGRAD_ALIGNED per layer shape.
The knobs are read once per process, so every configuration runs in its own subprocess.
usage: python tools/wgrad_sweep.py            (driver)      python tools/wgrad_sweep.py child   (one configuration)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [("l1s 64->144", 16, 64, 144, (32, 56, 56), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          ("l1t 144->64", 16, 144, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
          ("l2.0s 64->230 s2", 16, 64, 230, (32, 56, 56), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
          ("l2s 128->230", 16, 128, 230, (16, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          ("l2s 128->288", 16, 128, 288, (16, 28, 28), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          ("l2t 288->128", 16, 288, 128, (16, 28, 28), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
          ("l3s 256->576", 16, 256, 576, (8, 14, 14), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          ("l4s 512->1152", 16, 512, 1152, (4, 7, 7), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
          ("stem_t 45->64", 16, 45, 64, (32, 56, 56), (3, 1, 1), (1, 1, 1), (1, 0, 0))]


def child():
    import torch
    from selavi_b200 import ops
    from tools.quick_bench import timeit
    dev = torch.device("cuda:0")
    out = []
    for name, nb, ci, co, thw, k, s, p in LAYERS:
        geom = ops.ConvGeom(nb, ci, co, thw, k, s, p)
        x = torch.randn(geom.in_shape(), device=dev)
        dz = torch.randn(geom.out_shape(), device=dev)
        z_hi, z_lo = ops.split_bf16(dz)
        dw = torch.empty(co, ci, *k, device=dev)
        ms = timeit(lambda: ops.conv_wgrad_bf16(x, z_hi, z_lo, geom, dw), warm=2, rep=7)
        out.append(f"{ms:.3f}")
        del x, dz, z_hi, z_lo
    print("RESULT " + " ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        print("config (waves, aligned) | " + " | ".join(l[0] for l in LAYERS))
        for waves, aligned in [(1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1), (4, 1)]:
            env = dict(os.environ, SELAVI_WGRAD_WAVES=str(waves), SELAVI_WGRAD_ALIGNED=str(aligned))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(f"waves={waves} aligned={aligned}: " + (line[0][7:] if line else "FAILED " + r.stdout[-400:]), flush=True)
