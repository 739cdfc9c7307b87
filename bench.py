"""Benchmark of the SeLaVi data-parallel training hot path on B200 (contract: see the task statement / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W                # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...      # the reference's CPU path (oracle port) on host cores
    python bench.py --impl reference_gpu --gpus N --steps K ...  # the reference model on stock PyTorch + cuDNN + NCCL (GPU)

A step = one train step of BASELINE.json configs[1] (per-GPU batch 16 synthetic clips 3x32x112x112 + spectrograms
1x257x200, K=309, 10 heads): forward of both towers and the heads, 0.5*CE_v + 0.5*CE_a, zero_grad, backward
(DDP gradient all-reduce + SyncBN statistics for N > 1) and the SGD update.  `value` times it with the batch
resident in HBM; `e2e` repeats it through the public API from pinned host buffers (H2D of the batch and D2H of
the loss inside the timed region).  Next to it, in the same JSON line:
  sk        Sinkhorn-Knopp iterations/s on cfg-5's matrix (N=200000 x K=309 float64, rows sharded over the ranks, 100 iterations)
  sweep     eval-mode feature sweep (src/sk_utils.py:194-254) clips/s at the reference's sweep batch of 64 per GPU
  assign    heads + float64 softmax product + full Sinkhorn-Knopp solve of all 10 heads at VGG-Sound size (N=170752 rows)
  incl_sk   clips/s including the amortised sweep + assignment at the reference's schedule (nopts=100 over 100 epochs)
  fast_mode single-pass MMAs (SELAVI_MMA_PASSES=1) with the measured logit error beside it
  library_baseline  the reference model on stock PyTorch/cuDNN on the same GPU (TF32 convs allowed = torch default, and strict fp32)
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec (video+audio fwd/bwd + SK assign)"
CFG = dict(batch=16, T=32, HW=112, spec_T=200, K=309, hc=10)
# BASELINE.json configs: [1] (default, the configuration the metric is quoted on), [2] Kinetics-400 shape, [3] AVE shape with
# `match` and ind_groups=2; the train step differs only in K, the assignment phase in K, N and the head alignment
CONFIGS = {"cfg2": dict(idx=1, K=309, N=170752, match=False, ind_groups=1, name="VGG-Sound shape"),
           "cfg3": dict(idx=2, K=400, N=230976, match=False, ind_groups=1, name="Kinetics-400 shape"),
           "cfg4": dict(idx=3, K=28, N=3328, match=True, ind_groups=2, name="AVE shape, match=True, ind_groups=2")}
WORKLOAD = ("configs[1]: train step (video R(2+1)D-18 + audio ResNet-9 fwd/bwd, 20 MLP heads, CE, SGD), per-GPU batch 16, "
            "clips 3x32x112x112, spectrograms 1x257x200, K=309, 10 heads")


def select_config(name):
    global WORKLOAD, SK_DATASET_N
    c = CONFIGS[name]
    CFG["K"] = c["K"]
    SK_DATASET_N = c["N"]
    WORKLOAD = (f"configs[{c['idx']}] ({c['name']}): train step (video R(2+1)D-18 + audio ResNet-9 fwd/bwd, 20 MLP heads, CE, SGD), "
                f"per-GPU batch 16, clips 3x32x112x112, spectrograms 1x257x200, K={c['K']}, 10 heads")
    return c
# SURVEY §8d / BASELINE.md §3: algorithmic conv+linear FLOPs of one train step per sample (3x fwd - input dgrads)
FLOP_PER_SAMPLE_STEP = 490.3e9
# the reference's SK schedule (opt.py:71,88): nopts = 100 label optimisations over epochs = 100, i.e. on average one
# dataset-wide sweep + assignment per epoch; VGG-Sound (configs[1]) has N = 170752 clips
SK_DATASET_N, SK_NOPTS, SK_EPOCHS = 170752, 100, 100


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    except Exception:  # noqa: BLE001
        return dict(hbm=6650.0, tensor=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        return len(self.rows)

    def stop(self, first=0, last=None):
        """summary of the samples [first, last) (indices from mark()): the ones taken during the timed region"""
        if self.proc is None:
            return None
        self.proc.terminate()
        rows = [r for r in self.rows[first:last] if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return None
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def synthetic_batch(torch, rank, batch):
    g = torch.Generator().manual_seed(31 + rank)
    video = torch.randn(batch, 3, CFG["T"], CFG["HW"], CFG["HW"], generator=g)
    spec = torch.randn(batch, 1, 257, CFG["spec_T"], generator=g) * 17.89 + 1.93
    labels = torch.randint(0, CFG["K"], (batch, CFG["hc"]), generator=g)
    return video, spec, labels


# ------------------------------------------------------------------------------------------------ reference arms
def cpu_train_throughput(steps, warmup, batch, budget_s=None):
    """The reference's CPU path for this workload (oracle port: torchvision/torch.nn restatement of model.py +
    utils.get_loss + torch.optim.SGD, oracle/model_oracle.py — reproduces the real reference's outputs bit for bit,
    tests/test_oracle.py) on all host cores.  `budget_s`: stop after the step that exceeds it (the line reports the steps
    actually timed).  Returns (clips/s, cores, sample text, ms/step, steps timed)."""
    import torch
    from oracle.model_oracle import OracleAVModel, oracle_train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(31)
    model = OracleAVModel(CFG["hc"], CFG["K"]).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-5)
    video, spec, labels = synthetic_batch(torch, 0, batch)
    for _ in range(warmup):
        oracle_train_step(model, opt, video, spec, labels, CFG["hc"])
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        loss = oracle_train_step(model, opt, video, spec, labels, CFG["hc"])
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    float(loss.detach())
    dt = time.perf_counter() - t0
    sample = (f"{done} train step(s) of {batch} clips (configs[1] shapes 3x{CFG['T']}x{CFG['HW']}x{CFG['HW']} + "
              f"1x257x{CFG['spec_T']}, K={CFG['K']}, {CFG['hc']} heads) after {warmup} warm-up step(s), "
              f"torch {torch.__version__} CPU, {cores} threads")
    return batch * done / dt, cores, sample, dt / done * 1e3, done


def run_reference(args):
    """CPU arm: same config (batch 16, configs[1] shapes) and the same number of timed steps as our arm; the wall-clock
    budget (default 330 s of timed steps) only cuts the step count on a host too slow to finish them."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    val, cores, sample, ms, done = cpu_train_throughput(args.steps, warm, args.batch, budget_s=args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": done,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": args.batch, "parallelism": "cpu",
                       "note": "kind 'port': /root/reference does not exist on the GPU box; oracle/model_oracle.py reproduces the "
                               "unmodified reference bit for bit (tests/test_oracle.py, goldens generated from /root/reference)"},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_library_throughput(torch, dev, world, local, batch, steps, warmup, allow_tf32):
    """The reference model on STOCK PyTorch: torchvision/torch.nn modules (oracle/model_oracle.py, the bit-exact
    restatement of model.py) on CUDA through cuDNN / cuBLAS, torch.optim.SGD, cudnn.benchmark = True (main.py:187), for
    world > 1 SyncBatchNorm + DistributedDataParallel over NCCL (main.py:116-160).  `allow_tf32`: torch's default lets
    cuDNN convolutions use TF32; False = strict fp32 (the reference era's arithmetic).  Returns (clips/s, ms/step)."""
    import torch.distributed as dist
    from oracle.model_oracle import OracleAVModel, oracle_train_step
    old = torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = True
    try:
        torch.manual_seed(31)
        model = OracleAVModel(CFG["hc"], CFG["K"])
        if world > 1:
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = model.to(dev).train()
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-5)
        net = model
        if world > 1:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
        video, spec, labels = (x.to(dev) for x in synthetic_batch(torch, int(os.environ.get("RANK", "0")), batch))
        for _ in range(warmup):
            oracle_train_step(net, opt, video, spec, labels, CFG["hc"])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            oracle_train_step(net, opt, video, spec, labels, CFG["hc"])
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item()) / steps
        del net, model, opt
        torch.cuda.empty_cache()
        return batch * world / (ms * 1e-3), ms
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old


def run_reference_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for name, tf32 in (("tf32_convs_allowed (torch default)", True), ("fp32_strict", False)):
        v, ms = gpu_library_throughput(torch, dev, world, local, args.batch, args.steps, max(args.warmup, 5), tf32)
        out[name] = {"value": v, "unit": "clips/s", "ms_per_step": ms}
    if rank == 0:
        best = out["tf32_convs_allowed (torch default)"]
        line = {"impl": "reference_gpu", "metric": METRIC, "value": best["value"], "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 5), "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 storage; cuDNN convolutions may use TF32 (torch default)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": args.batch * world, "parallelism": f"dp{world}"},
                "library_baseline": out}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from selavi_b200 import _lib, build, engine, model as sv_model, ops
    from selavi_b200.optim import SGD
    from selavi_b200.sk_utils import SKComm, SKWorkspace, sk_solve_raw, softmax_product
    from selavi_b200.utils import get_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=480))   # a hang fails in minutes, not tens of minutes
        dist.barrier()
    _lib.lib()
    pk = peaks()
    B, hc, K = args.batch, CFG["hc"], CFG["K"]

    torch.manual_seed(31)
    with contextlib.redirect_stdout(sys.stderr):       # the reference's constructor prints; stdout carries ONE JSON line
        model = sv_model.load_model(vid_base_arch="r2plus1d_18", aud_base_arch="resnet9", pretrained=False,
                                    norm_feat=False, use_mlp=True, headcount=hc, num_classes=K)
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)           # main.py:117-118
    model = model.to(dev).train()
    opt = SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-5)    # main.py:132-137
    net = model
    if world > 1:
        ddp_kw = {}
        if os.environ.get("SELAVI_DDP_BUCKET_MB"):   # (experiment knob; default = the reference's call, main.py:156-160)
            ddp_kw["bucket_cap_mb"] = int(os.environ["SELAVI_DDP_BUCKET_MB"])
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True, **ddp_kw)
    video_h, spec_h, labels_h = synthetic_batch(torch, rank, B)
    video_h, spec_h, labels_h = video_h.pin_memory(), spec_h.pin_memory(), labels_h.pin_memory()
    video_d, spec_d, labels_d = video_h.to(dev), spec_h.to(dev), labels_h.to(dev)

    # fast mode (single-pass MMAs): error of the 512-d video features against parity mode on the FRESH weights (train-mode
    # BatchNorm, same batch).  Features, not logits: the bench batch is 16 i.i.d. white-noise clips whose features are
    # nearly identical, and the heads' BatchNorm1d over such a batch amplifies ANY rounding difference ~100x (for logits on
    # a well-conditioned batch see profiles/r02_precision_modes.txt: 1.4e-2 fast vs 6.0e-5 parity at this shape).
    fast_err = None
    if world > 1:
        args.no_fast_mode = True      # a single-GPU numerics reference point; not repeated under data parallelism
    if not args.no_fast_mode:
        try:
            model.return_features = True
            with torch.no_grad():
                f3, _ = net(video_d, spec_d)
                old_passes, engine.PASSES = engine.PASSES, 1
                try:
                    f1, _ = net(video_d, spec_d)
                finally:
                    engine.PASSES = old_passes
                fast_err = float((f1.double() - f3.double()).norm() / f3.double().norm())
            del f3, f1
        except Exception as e:  # noqa: BLE001
            fast_err = repr(e)[:200]
        finally:
            model.return_features = False

    def train_step(video, spec, labels):
        fv, fa = net(video, spec)
        loss = 0.5 * get_loss(fv, labels, hc) + 0.5 * get_loss(fa, labels, hc)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_step=None):
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        barrier()
        if per_step is not None:
            per_step.extend(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        ms = torch.tensor([evs[0].elapsed_time(evs[steps])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # warm-up: the W requested steps, then (untimed) until the step time has settled -- on a fresh box the first steps
    # also pay for the caching allocator growing to its 16 GB working set, first-use module loads and clock ramp-up
    # the clock sampler (an nvidia-smi child process) starts BEFORE the warm-up: its NVML initialisation briefly blocks
    # the driver and used to land inside the timed region (sporadic 90-130 ms steps right after its start)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    warm_ms = []
    for _ in range(args.warmup):
        timed(lambda: train_step(video_d, spec_d, labels_d), 1, warm_ms)
    while len(warm_ms) < args.warmup + 12 and not os.environ.get("SELAVI_BENCH_NO_SETTLE"):   # (off under ncu)
        settled = torch.tensor([1.0 if abs(warm_ms[-1] - warm_ms[-2]) < 0.03 * warm_ms[-1] and
                                abs(warm_ms[-2] - warm_ms[-3]) < 0.03 * warm_ms[-1] else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(settled, op=dist.ReduceOp.MIN)
        if settled.item() > 0:
            break
        timed(lambda: train_step(video_d, spec_d, labels_d), 1, warm_ms)
    _lib.COUNT_CALLS = True
    _lib.CALLS.clear()
    step_ms = []
    mark0 = sampler.mark() if sampler else 0
    ms_total = timed(lambda: train_step(video_d, spec_d, labels_d), args.steps, step_ms)
    launches = _lib.kernel_launches()
    _lib.COUNT_CALLS = False
    # a timed region disturbed from outside (one step > 1.5x the settled warm-up step) is repeated ONCE; the first
    # attempt stays in the JSON line (`first_attempt_step_ms`)
    first_attempt = None
    disturbed = torch.tensor([1.0 if max(step_ms) > 1.5 * min(warm_ms[-3:]) else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(disturbed, op=dist.ReduceOp.MAX)
    if disturbed.item() > 0:
        first_attempt, step_ms = step_ms, []
        mark0 = sampler.mark() if sampler else 0
        ms_total = timed(lambda: train_step(video_d, spec_d, labels_d), args.steps, step_ms)
    # host-side enqueue time of one step (python + ctypes + torch allocator), GPU idle at the start
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train_step(video_d, spec_d, labels_d)
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    # end to end through the public API from pinned host buffers: every step copies ITS batch host->device (on a copy
    # stream, issued while the previous step computes, as a prefetching data loader would) and reads the loss back
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [[torch.empty_like(video_d), torch.empty_like(spec_d), torch.empty_like(labels_d), torch.cuda.Event()] for _ in range(2)]
    state = {"k": 0}

    def h2d(slot):
        copy_stream.wait_stream(torch.cuda.current_stream(dev))     # the slot's previous consumer has been enqueued
        with torch.cuda.stream(copy_stream):
            slot[0].copy_(video_h, non_blocking=True)
            slot[1].copy_(spec_h, non_blocking=True)
            slot[2].copy_(labels_h, non_blocking=True)
            slot[3].record(copy_stream)

    def e2e_step():
        cur = slots[state["k"] & 1]
        nxt = slots[(state["k"] + 1) & 1]
        state["k"] += 1
        torch.cuda.current_stream(dev).wait_event(cur[3])
        loss = train_step(cur[0], cur[1], cur[2])
        h2d(nxt)                                                      # next step's batch: overlaps this step's kernels
        return float(loss.item())

    h2d(slots[0])
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop(mark0) if sampler else None   # samples taken during the two timed regions (device-resident, e2e)
    del slots

    # ---- live per-kernel timing of one more step (CUDA events around every conv launch on the launching stream)
    ops.PROFILE = []
    side, engine.WGRAD_STREAM = engine.WGRAD_STREAM, False   # per-kernel times: no concurrent weight-gradient stream
    aud, engine.AUDIO_STREAM = engine.AUDIO_STREAM, False    # ... and no concurrent audio-tower stream
    train_step(video_d, spec_d, labels_d)
    torch.cuda.synchronize()
    engine.WGRAD_STREAM, engine.AUDIO_STREAM = side, aud
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    if os.environ.get("SELAVI_BENCH_DETAIL") and rank == 0:
        for kind, flops, e0, e1, tag, kern in prof:
            ms = e0.elapsed_time(e1)
            print(f"DETAIL {kind:11s} ci,co,T,H,W,k,s={tag} {ms:8.3f} ms {flops / ms / 1e9:7.1f} TF/s  {kern}", file=sys.stderr)
    for _kind, flops, e0, e1, _tag, kern in prof:
        a = agg.setdefault(kern, [0.0, 0.0, 0])
        a[0] += flops
        a[1] += e0.elapsed_time(e1)
        a[2] += 1
    ms_step = ms_total / args.steps
    # per kernel: algorithmic conv FLOPs (2*M*Co*Ci*taps) / launch time, CUDA events on the launching
    # stream.  The x3 operand split issues 3 MMAs per algorithmic MAC, so `issued_mma_tflops` = 3x achieved; the fp16x3 /
    # bf16x3 kernels run at the bf16 rate (ceiling 1/3 of the peak below), the tf32x3 kernels at half of it (ceiling 1/6).
    kernels = {}
    for kern, (fl, ms, n) in agg.items():
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        kernels[kern] = {"achieved": ach, "unit": "TFLOP/s", "frac": ach / pk["tensor"], "issued_mma_tflops": ach * engine.PASSES,
                         "launches_per_step": n, "avg_launch_ms": ms / max(n, 1), "ms_per_step": ms, "share_of_step": ms / ms_step,
                         "algorithmic_flops_per_launch": fl / max(n, 1)}
    top = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    conv_ms = sum(k["ms_per_step"] for k in kernels.values())
    conv_fl = sum(a[0] for a in agg.values())
    roofline = {"kernel": top, "bound": "tensor", "achieved": kernels[top]["achieved"] if top else 0.0,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": kernels[top]["frac"] if top else 0.0, "traffic": None,
                "peak_source": f"{pk['src']} bf16 dense sustained", "launches_per_step": kernels[top]["launches_per_step"] if top else 0,
                "avg_launch_ms": kernels[top]["avg_launch_ms"] if top else 0.0,
                "share_of_step": kernels[top]["share_of_step"] if top else 0.0,
                "mma_passes": engine.PASSES, "issued_mma_tflops": kernels[top]["issued_mma_tflops"] if top else 0.0,
                "all_conv_kernels": {"achieved": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0, "unit": "TFLOP/s",
                                     "share_of_step": conv_ms / ms_step},
                "per_kernel": kernels}

    # DRAM bytes per launch of the reported kernel: ncu `--set full` capture of ONE train step of this same command
    # (tools/ncu_traffic.py -> profiles/ncu_traffic.json): dram__bytes_read.sum + dram__bytes_write.sum summed over every
    # launch of the kernel set the timing above brackets (e.g. wgrad + split + reduce), divided by the bracketed launches
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except (OSError, ValueError):
        pass
    if top and isinstance(traffic.get(top), dict):
        roofline["traffic"] = traffic[top].get("bytes_per_launch")
        roofline["traffic_source"] = traffic[top].get("source")

    # ---- Sinkhorn-Knopp: cfg-5 matrix, rows sharded over ranks, exactly 100 iterations (convergence test computed)
    sk = None
    comms = {}
    try:
        N, iters = 200000, 100
        n_local = N // world
        g = torch.Generator(device=dev).manual_seed(rank)
        PS = torch.softmax(torch.randn(n_local, K, dtype=torch.float64, device=dev, generator=g), 1) * \
            torch.softmax(torch.randn(n_local, K, dtype=torch.float64, device=dev, generator=g), 1)
        ws = SKWorkspace(K, n_local, dev)
        comm = comms.setdefault(K, SKComm(K)) if world > 1 else None
        kw = dict(world=world, rank=rank, peer_sum=comm.sum.peer_ptrs, peer_flag=comm.flag.peer_ptrs) if comm else {}

        def sk_run(prep, n):
            if comm:
                comm.reset()
            sk_solve_raw(PS, n_local * world, 20.0, None, ws, max_iters=n, stop_on_converge=False, do_prep=prep,
                         do_final=False, **kw)

        sk_run(True, 10)
        for _ in range(3):
            sk_run(False, iters)
        ms_sk = min(timed(lambda: sk_run(False, iters), 1) for _ in range(3))
        bytes_iter = n_local * K * 8
        gbs = bytes_iter * iters / (ms_sk * 1e-3) / 1e9
        sk = {"iters_per_sec": iters / (ms_sk * 1e-3), "N": n_local * world, "K": K, "iters": iters, "rows_per_gpu": n_local,
              "roofline": {"kernel": "sk_kernel", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                           "frac": gbs / pk["hbm"], "peak_source": pk["src"],
                           "traffic": (traffic.get("sk_kernel") or {}).get("bytes_per_launch") if world == 1 else None,
                           "algorithmic_bytes_per_launch": bytes_iter * iters,
                           "note": "per GPU; shards below ~126 MB are L2-resident, so frac can exceed 1"}}
        del PS, ws
    except Exception as e:  # noqa: BLE001
        sk = {"error": repr(e)[:300]}

    # ---- eval-mode feature sweep (src/sk_utils.py:194-254): both towers, running-stat BN folded into the conv loaders,
    #      return_features, batch 64 per GPU like the reference's sweep DataLoader
    sweep = assign = incl = None
    try:
        SB = 64
        gsw = torch.Generator().manual_seed(77 + rank)
        v64 = torch.randn(SB, 3, CFG["T"], CFG["HW"], CFG["HW"], generator=gsw).to(dev)
        s64 = (torch.randn(SB, 1, 257, CFG["spec_T"], generator=gsw) * 17.89 + 1.93).to(dev)
        model.eval()
        model.return_features = True

        def sweep_step():
            with torch.no_grad():
                return net(v64, s64)

        for _ in range(2):
            sweep_step()
        nsw = 5
        ms_sw = timed(sweep_step, nsw) / nsw
        model.return_features = False
        model.train()
        sweep_cps = SB * world / (ms_sw * 1e-3)
        sweep = {"value": sweep_cps, "unit": "clips/s", "ms_per_batch": ms_sw, "batch_per_gpu": SB, "steps": nsw,
                 "algorithmic_tflops": 163.9e9 * SB / (ms_sw * 1e-3) / 1e12,
                 "what": "eval-mode forward of both towers -> 512-d features (get_cluster_assignments_gpu sweep body)"}
        del v64, s64
        # ---- assignment at VGG-Sound size: per head  heads(F_v), heads(F_a) -> float64 softmax product -> full SK solve
        n_rows = SK_DATASET_N // world
        gf = torch.Generator(device=dev).manual_seed(5 + rank)
        F_v = torch.randn(n_rows, 512, device=dev, generator=gf).abs()
        F_a = torch.randn(n_rows, 512, device=dev, generator=gf).abs()
        model.eval()
        ws2 = SKWorkspace(K, n_rows, dev)
        comm = comms.setdefault(K, SKComm(K)) if world > 1 else None
        kw = dict(world=world, rank=rank, peer_sum=comm.sum.peer_ptrs, peer_flag=comm.flag.peer_ptrs) if comm else {}
        iters_seen = []

        match_s = [0.0]

        def assign_all():
            with torch.no_grad():
                for h in range(hc):
                    lv = getattr(model, f"mlp_v{h}").forward(F_v)
                    la = getattr(model, f"mlp_a{h}").forward(F_a)
                    if args.cfg["match"]:      # first SK call of cfg-4: align the audio head to the video head first
                        import types
                        from selavi_b200.sk_utils import match_order
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        match_order(types.SimpleNamespace(rank=rank), lv, la, list(getattr(model, f"mlp_a{h}").modules())[-1], logits=True)
                        torch.cuda.synchronize()
                        match_s[0] += time.perf_counter() - t0
                        la = getattr(model, f"mlp_a{h}").forward(F_a)
                    PSh = softmax_product(lv, la)
                    if comm:
                        comm.reset()
                    sk_solve_raw(PSh, n_rows * world, 20.0, None, ws2, **kw)
            iters_seen.append(int(ws2.iters.item()))

        assign_all()
        match_s[0] = 0.0
        ms_as = timed(assign_all, 1)
        model.train()
        assign = {"seconds": ms_as * 1e-3, "heads": hc, "rows": n_rows * world, "K": K, "sk_iters_last_head": iters_seen[-1],
                  **({"match_order_seconds": match_s[0], "ind_groups": args.cfg["ind_groups"],
                      "note": "match_order = K x K L1 cost-matrix kernel + host hill-climb (np.random stream of the reference); with "
                              "ind_groups=2 the reference sweeps the dataset twice (see 'incl_sk')"} if args.cfg["match"] else {}),
                  "what": "10 x (2 MLP heads on [N,512] features, float64 softmax product, Sinkhorn-Knopp to convergence, argmax)"}
        del F_v, F_a, ws2
        train_cps = B * world / (ms_step * 1e-3)
        per_clip = 1.0 / train_cps + (SK_NOPTS / SK_EPOCHS) * (args.cfg["ind_groups"] / sweep_cps + ms_as * 1e-3 / SK_DATASET_N)
        incl = {"value": 1.0 / per_clip, "unit": "clips/s",
                "formula": f"1 / (1/train + (nopts/epochs) * (ind_groups/sweep + assign_seconds/N)), nopts=100, epochs=100, "
                           f"N={SK_DATASET_N}, ind_groups={args.cfg['ind_groups']} (opt.py:71,88,102)"}
    except Exception as e:  # noqa: BLE001
        sweep = sweep or {"error": repr(e)[:300]}
        model.return_features = False
        model.train()

    # ---- fast mode: single-pass MMAs (SELAVI_MMA_PASSES=1; every conv on the tf32 implicit-GEMM kernels)
    fast = None
    if not args.no_fast_mode:
        try:
            old_passes, engine.PASSES = engine.PASSES, 1
            try:
                for _ in range(3):
                    train_step(video_d, spec_d, labels_d)
                nf = max(3, min(args.steps, 10))
                ms_fast = timed(lambda: train_step(video_d, spec_d, labels_d), nf) / nf
            finally:
                engine.PASSES = old_passes
            fast = {"value": B * world / (ms_fast * 1e-3), "unit": "clips/s", "ms_per_step": ms_fast, "steps": nf,
                    "video_feature_rel_err_vs_parity_mode": fast_err,
                    "note": "single-pass tf32 MMAs on the implicit-GEMM kernels only (the tap-reuse fp16x3 kernels have no single-pass "
                            "variant, so this mode is SLOWER than parity mode; train-mode logits 1.4e-2 off, "
                            "profiles/r02_precision_modes.txt): kept as a numerics reference point, not as a product mode"}
        except Exception as e:  # noqa: BLE001
            fast = {"error": repr(e)[:300]}

    # ---- the reference model on stock PyTorch + cuDNN on this GPU (the "library" bar of BASELINE.md §4)
    library = None
    if (world == 1 and not args.no_library_baseline) or args.library_baseline:
        try:
            del net, opt, model
            torch.cuda.empty_cache()
            library = {}
            for name, tf32 in (("tf32_convs_allowed (torch default)", True), ("fp32_strict", False)):
                v, ms = gpu_library_throughput(torch, dev, world, local, B, max(3, min(args.steps, 10)), 5, tf32)
                library[name] = {"value": v, "unit": "clips/s", "ms_per_step": ms}
            library["what"] = ("torchvision r2plus1d_18 + ResNet-9 + MLP heads (oracle/model_oracle.py) on CUDA via cuDNN/cuBLAS, "
                               "torch.optim.SGD, cudnn.benchmark=True, same batch; also `bench.py --impl reference_gpu`")
        except Exception as e:  # noqa: BLE001
            library = {"error": repr(e)[:300]}

    if rank == 0:
        global_batch = B * world
        value = global_batch / (ms_step * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": len(warm_ms), "ms_per_step": ms_step, "step_ms": [round(x, 2) for x in step_ms],
                "warmup_step_ms": [round(x, 2) for x in warm_ms],
                **({"first_attempt_step_ms": [round(x, 2) for x in first_attempt]} if first_attempt else {}), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (operands split hi/lo: fp16x3 / tf32x3 forward, bf16x3 backward MMAs; fp32 accumulate and storage)" if engine.PASSES == 3 else "tf32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "global_batch": global_batch, "parallelism": f"dp{world}", "l2": "inputs_exceed_l2 (16 GB of activations per step)",
                           "value_is": "train step only; Sinkhorn-Knopp and the feature sweep are timed separately (keys 'sk', 'sweep', "
                                       "'assign') and folded in at the reference's schedule in 'incl_sk'"},
                "e2e": {"value": global_batch / (ms_e2e / args.steps * 1e-3), "unit": "clips/s",
                        "h2d_bytes_per_step": int(video_h.numel() * 4 + spec_h.numel() * 4 + labels_h.numel() * 8),
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "gpu_launches_per_step": launches // args.steps, "host_enqueue_ms_per_step": host_ms, "roofline": roofline, "sk": sk,
                "sweep": sweep, "assign": assign, "incl_sk": incl, "fast_mode": fast, "library_baseline": library, "clocks": clocks,
                "algorithmic_tflops": FLOP_PER_SAMPLE_STEP * B / (ms_step * 1e-3) / 1e12}
        if world == 1 and not args.no_cpu_baseline:
            val, cores, sample, _, _ = cpu_train_throughput(2, 1, B, budget_s=40)
            line["cpu_baseline"] = {"value": val, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=CFG["batch"], help="per-GPU batch (configs[1]: 16)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_gpu"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configs[1] (default) / [2] / [3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--library-baseline", action="store_true", help="also at N > 1 (default: N = 1 only)")
    ap.add_argument("--no-fast-mode", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=330.0, help="--impl reference: wall-clock cap of the timed steps")
    args = ap.parse_args()
    args.cfg = select_config(args.config)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_gpu":
        run_reference_gpu(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
