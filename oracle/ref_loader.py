"""Load the UNMODIFIED reference (read-only /root/reference) for pinning the oracle.  TEST INFRASTRUCTURE.

Only usable in the authoring container (the GPU box has no /root/reference).  Two arithmetic-neutral shims:
  * torchvision 0.4's private `_resnet(arch, block, layers, pretrained, progress)` signature used at
    model.py:114 is restored;
  * src/sk_utils.py hard-codes device='cuda' (sk_utils.py:366-413); for a CPU run the source TEXT is loaded
    and those device spellings are replaced before exec — no reference code is copied into this repo.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("SELAVI_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(REF) and os.path.exists(os.path.join(REF, "model.py"))


def _shim_torchvision():
    import torchvision
    from torchvision.models.resnet import ResNet
    torchvision.models.resnet._resnet = (
        lambda arch, block, layers, pretrained, progress, **kw: ResNet(block, layers, **kw))


def load_model_module():
    """reference model.py as a module named `ref_model`."""
    _shim_torchvision()
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_utils_get_loss():
    """reference utils.get_loss (utils.py:377-387) without importing utils' heavy dependencies."""
    src = open(os.path.join(REF, "utils.py")).read()
    start = src.index("def get_loss(")
    end = src.index("def warmup_batchnorm(")
    mod = types.ModuleType("ref_utils_get_loss")
    import torch
    mod.torch = torch
    exec(compile(src[start:end], os.path.join(REF, "utils.py"), "exec"), mod.__dict__)
    return mod.get_loss


def load_sk_module(device="cpu", sweep=False):
    """reference src/sk_utils.py with its hard-coded cuda device strings substituted (text-level).
    sweep=True additionally neutralises the CUDA-only spellings of `get_cluster_assignments_gpu` / `match_order`
    (non-blocking copies, torch.cuda tensor types, synchronize / empty_cache, `.to('cuda')`) so that the whole sweep +
    assignment bookkeeping runs on the CPU under a single-process gloo group; no arithmetic or control flow changes."""
    src = open(os.path.join(REF, "src", "sk_utils.py")).read()
    if device == "cpu":
        src = src.replace("device='cuda:0'", "device='cpu'").replace("device='cuda'", "device='cpu'")
        if sweep:
            src = src.replace(".cuda(non_blocking=True)", "")
            src = src.replace("torch.cuda.DoubleTensor", "torch.DoubleTensor").replace("torch.cuda.FloatTensor", "torch.FloatTensor")
            src = src.replace("torch.cuda.synchronize()", "None").replace("torch.cuda.empty_cache()", "None")
            src = src.replace(".to('cuda')", "")
        src = src.replace(".cuda()", "")
    mod = types.ModuleType("ref_sk_utils")
    mod.__file__ = os.path.join(REF, "src", "sk_utils.py")
    sys.path.insert(0, REF)
    try:
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    finally:
        sys.path.remove(REF)
    return mod
