"""B200-native mirror of the hot-path pieces of the reference `utils.py`: `get_loss` (utils.py:377-387) on the
fused cross-entropy kernel, and `warmup_batchnorm` (utils.py:389-418)."""
import time

import torch
import torch.distributed as dist

from .engine import get_loss  # noqa: F401

__all__ = ["get_loss", "warmup_batchnorm"]


def warmup_batchnorm(args, model, dataloader, batches=20, group=None):
    """Same contract as utils.py:389-418 (train-mode no-grad forwards to warm the BN running statistics)."""
    print("Warming up batchnorm", flush=True)
    start = time.time()
    with torch.no_grad():
        model.train()
        for i, batch in enumerate(dataloader):
            video, audio, _, _, idx = batch
            video = video.cuda(non_blocking=True)
            audio = audio.cuda(non_blocking=True)
            if i == batches:
                break
            _ = model(video, audio)
        if getattr(args, "distributed", False) and args.world_size > 1:
            if group is not None:
                dist.barrier(group=group)
            else:
                dist.barrier()
    print(f"Finshed warming up batchnorm!) took {(time.time()-start)/60:.1f}min", flush=True)
