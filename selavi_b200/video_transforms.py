"""B200-native mirror of the reference `datasets/video_transforms.py:clip_augmentation` (:462-504, default flags).

The reference runs this per clip on the CPU inside DataLoader workers (float conversion, normalisation, bilinear
short-side scale jitter, crop, flip, layout change: ~6 full passes over a 32-frame clip).  Here the decoded uint8
frames go to the GPU as they are (4x fewer H2D bytes than float32) and the whole chain is ONE kernel
(csrc/video.cu).  The random draws stay on the host and follow the reference's np.random call order exactly
(`draw_params`), so a seeded run picks the same scale / crop / flip as the reference.

`clip_augmentation` keeps the reference signature for one clip; `clip_augmentation_batch` is the batched entry point
for the training process.  No CPU fallback: frames must be a CUDA uint8 tensor.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib


def _scaled_size(height, width, size):
    # random_short_side_scale_jitter, video_transforms.py:54-68
    if (width <= height and width == size) or (height <= width and height == size):
        return height, width
    if width < height:
        return int(math.floor((float(height) / width) * size)), size
    return size, int(math.floor((float(width) / height) * size))


def draw_params(height, width, spatial_idx=-1, min_scale=256, max_scale=320, crop_size=224, rng=np.random):
    """The random draws of spatial_sampling (video_transforms.py:420-459) in the reference's order
    -> (new_h, new_w, y_off, x_off, flip)."""
    if spatial_idx not in (-1, 0, 1, 2, 3, 4, 5):
        raise AssertionError("spatial_idx must be in [-1, 0, 1, 2, 3, 4, 5]")
    size = int(round(rng.uniform(min_scale, max_scale)))
    nh, nw = _scaled_size(height, width, size)
    if spatial_idx == -1:
        y = x = 0
        if not (nh == crop_size and nw == crop_size):
            if nh > crop_size:
                y = int(rng.randint(0, nh - crop_size))
            if nw > crop_size:
                x = int(rng.randint(0, nw - crop_size))
        flip = bool(rng.uniform() < 0.5)
    else:
        idx = spatial_idx % 3
        y = int(math.ceil((nh - crop_size) / 2))
        x = int(math.ceil((nw - crop_size) / 2))
        if nh > nw:
            y = 0 if idx == 0 else (nh - crop_size if idx == 2 else y)
        else:
            x = 0 if idx == 0 else (nw - crop_size if idx == 2 else x)
        flip = spatial_idx in (3, 4, 5)
        if flip:
            rng.uniform()   # horizontal_flip(1, .) still consumes one draw (video_transforms.py:157)
    return nh, nw, y, x, flip


def clip_augmentation_batch(frames, params=None, spatial_idx=-1, min_scale=256, max_scale=320, crop_size=224, rng=np.random,
                            out=None):
    """frames: CUDA uint8 [N, T, H, W, 3]; params: optional list of N (new_h, new_w, y_off, x_off, flip) tuples (drawn with
    `draw_params` clip by clip when omitted) -> float32 [N, 3, T, crop, crop] on the same device (current stream)."""
    if not (torch.is_tensor(frames) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 5 and frames.shape[-1] == 3):
        raise ValueError("clip_augmentation_batch needs a CUDA uint8 tensor [N, T, H, W, 3] (no CPU fallback)")
    frames = frames.contiguous()
    n, t, h, w, _ = frames.shape
    if params is None:
        params = [draw_params(h, w, spatial_idx, min_scale, max_scale, crop_size, rng) for _ in range(n)]
    if len(params) != n:
        raise ValueError("one parameter tuple per clip")
    arr = (ctypes.c_int * (5 * n))(*[int(v) for q in params for v in q])
    if out is None:
        out = torch.empty((n, 3, t, crop_size, crop_size), dtype=torch.float32, device=frames.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (n, 3, t, crop_size, crop_size)):
        raise ValueError("out must be a contiguous float32 CUDA tensor [N, 3, T, crop, crop]")
    with torch.cuda.device(frames.device):
        _lib.check(_lib.lib().selavi_clip_augment(_lib.ptr(frames), _lib.ptr(out), n, t, h, w, crop_size, arr, _lib.stream_ptr()),
                   "selavi_clip_augment")
    return out


def clip_augmentation(frames, spatial_idx=-1, min_scale=256, max_scale=320, crop_size=224, colorjitter=False,
                      use_grayscale=False, use_gaussian=False):
    """Reference signature (video_transforms.py:462-471) for ONE clip: uint8 [T, H, W, 3] CUDA -> float32 [3, T, crop, crop].
    colour jitter / grayscale (off by default, opt.py:47-50) are outside this path."""
    if colorjitter or use_grayscale:
        raise NotImplementedError("colorjitter / use_grayscale are not part of the default training path")
    return clip_augmentation_batch(frames[None], None, spatial_idx, min_scale, max_scale, crop_size)[0]
