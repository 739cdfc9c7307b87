"""Host-side wall time of each phase of one train step (no synchronisation inside the step, GPU idle at the start):
finds host-blocking calls in the multi-GPU path.  torchrun --nproc-per-node N tools/host_phase_times.py"""
import cProfile
import contextlib
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CFG, synthetic_batch  # noqa: E402
from selavi_b200 import model as sv_model  # noqa: E402
from selavi_b200.optim import SGD  # noqa: E402
from selavi_b200.utils import get_loss  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hc, K = CFG["hc"], CFG["K"]
    torch.manual_seed(31)
    with contextlib.redirect_stdout(sys.stderr):
        model = sv_model.load_model(vid_base_arch="r2plus1d_18", aud_base_arch="resnet9", pretrained=False, norm_feat=False,
                                    use_mlp=True, headcount=hc, num_classes=K)
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev).train()
    opt = SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-5)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    video, spec, labels = (t.to(dev) for t in synthetic_batch(torch, rank, CFG["batch"]))

    def step(times=None):
        t = [time.perf_counter()]
        fv, fa = net(video, spec)
        t.append(time.perf_counter())
        loss = 0.5 * get_loss(fv, labels, hc) + 0.5 * get_loss(fa, labels, hc)
        opt.zero_grad()
        t.append(time.perf_counter())
        loss.backward()
        t.append(time.perf_counter())
        opt.step()
        t.append(time.perf_counter())
        if times is not None:
            times.append([(b - a) * 1e3 for a, b in zip(t, t[1:])])

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        step(times)
        torch.cuda.synchronize()
    if rank == 0:
        for t in times:
            print("host ms  forward %.1f | loss+zero_grad %.1f | backward %.1f | opt.step %.1f" % tuple(t), flush=True)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    torch.cuda.synchronize()
    if rank == 0:
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    # GPU-side kernel times of one step from CUPTI (torch.profiler), rank 0
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        if rank == 0:
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60), flush=True)
    except Exception as e:  # noqa: BLE001
        print("profiler unavailable:", repr(e)[:200])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
