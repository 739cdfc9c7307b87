"""Multi-GPU parity (DDP + SyncBN train step, row-sharded Sinkhorn-Knopp, row-sharded sweep) on 2 / 4 / 8 B200s of one node:
spawns tests/mgpu_worker.py under torchrun and checks its verdict.  Each case is skipped when fewer GPUs are visible.  The
float64 CPU oracle of the DDP step is computed HERE once (oracle/model_oracle.py, full batch in one process — which is what
SyncBatchNorm + gradient averaging compute) and handed to the workers as an .npz."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _oracle_npz(path):
    import mgpu_worker as w
    from oracle.model_oracle import OracleAVModel, oracle_get_loss
    video, spec, labels = w.ddp_inputs()
    out = {}
    grads = {}
    for dtype in (torch.float32, torch.float64):
        m = w.build_model(lambda: OracleAVModel(w.HC, w.K)).to(dtype).train()
        fv, fa = m(torch.from_numpy(video).to(dtype), torch.from_numpy(spec).to(dtype))
        lab = torch.from_numpy(labels)
        loss = 0.5 * oracle_get_loss(fv, lab, w.HC) + 0.5 * oracle_get_loss(fa, lab, w.HC)
        loss.backward()
        grads[dtype] = {n: p.grad.detach().double().numpy() for n, p in m.named_parameters()}
        out["loss64" if dtype == torch.float64 else "loss32"] = float(loss)
    num = den = 0.0
    for n, g64 in grads[torch.float64].items():
        out["grad64/" + n] = g64
        out["err32/" + n] = np.linalg.norm(grads[torch.float32][n] - g64) / (np.linalg.norm(g64) + 1e-30)
        num += float(((grads[torch.float32][n] - g64) ** 2).sum())
        den += float((g64 ** 2).sum())
    out["global_err32"] = (num / den) ** 0.5     # the fp32 reference's own error over the concatenated gradient
    np.savez(path, **out)


@pytest.fixture(scope="module")
def oracle_file(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("mgpu") / "ddp_oracle.npz")
    _oracle_npz(path)
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(cuda_device, request, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} visible GPUs, found {torch.cuda.device_count()}")
    oracle_file = request.getfixturevalue("oracle_file")      # (the float64 CPU oracle is only computed when a case runs)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "mgpu_worker.py"), oracle_file]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    print(res.stdout[-6000:])
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"mgpu_worker_n{world}.log"), "w") as f:
            f.write(res.stdout)
    assert res.returncode == 0 and "MGPU_CHECK PASS" in res.stdout
