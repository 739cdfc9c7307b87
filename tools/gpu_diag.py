"""Run each GPU check in its own subprocess under a timeout (a hung kernel must not take the box down) and
print one line per check.  Usage on the GPU box:  python tools/gpu_diag.py [pattern]"""
import subprocess
import sys
import time

CHECKS = [
    ("sk_small", "tests/test_sk_gpu.py::test_sk_matches_oracle"),
    ("sk_golden", "tests/test_sk_gpu.py::test_sk_matches_reference_golden"),
    ("sk_cont", "tests/test_sk_gpu.py::test_sk_fixed_iterations_and_continuation"),
    ("sk_full", "tests/test_sk_gpu.py::test_sk_full_size_properties"),
    ("conv_fwd_all", "tests/test_conv_gpu.py::test_conv_forward"),
    ("conv_fwd_pro", "tests/test_conv_gpu.py::test_conv_forward_fused_bn_relu_prologue_and_stats"),
    ("conv_dgrad", "tests/test_conv_gpu.py::test_conv_dgrad"),
    ("conv_wgrad", "tests/test_conv_gpu.py::test_conv_wgrad"),
    ("conv_bf16bwd", "tests/test_conv_gpu.py::test_conv_dgrad_bf16x3"),
    ("heads", "tests/test_model_gpu.py::test_heads_linear_and_dropout_paths"),
    ("ce", "tests/test_model_gpu.py::test_ce_loss_matches_torch"),
    ("sgd", "tests/test_model_gpu.py::test_sgd_matches_torch"),
    ("model_train", "tests/test_model_gpu.py::test_train_step_matches_reference"),
    ("model_eval", "tests/test_model_gpu.py::test_eval_features_match_reference"),
]

if __name__ == "__main__":
    pats = sys.argv[1:]
    for name, target in CHECKS:
        if pats and not any(p in name for p in pats):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", target, "-q", "-s", "--no-header", "-p", "no:cacheprovider"],
                               capture_output=True, text=True, timeout=240)
            lines = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in lines if (("rel=" in l or "rel err" in l or "worst" in l) and "print" not in l) or l.startswith("FAILED") or l.startswith("E  ")]
            tail = "\n".join(keep[:60] + lines[-2:])
            status = "PASS" if r.returncode == 0 else f"FAIL({r.returncode})"
        except subprocess.TimeoutExpired as e:
            status, tail = "TIMEOUT", ((e.stdout or b"").decode(errors="replace")[-2000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-2000:])
        print(f"==== {name}: {status} in {time.time() - t0:.1f}s\n{tail}\n", flush=True)
