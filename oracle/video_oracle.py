"""CPU oracle of the GPU clip augmentation (uint8 THWC frames -> normalised, scale-jittered, cropped, flipped CTHW
float32 clip).  TEST INFRASTRUCTURE ONLY (imported by tests/ only).

Restates datasets/video_transforms.py of the reference with the random draws made explicit:
  clip_augmentation (:462-504)  float, /255, -MEAN, /STD (:474-477), THWC -> TCHW (:480), spatial_sampling, -> CTHW (:503)
  spatial_sampling (:420-459)   spatial_idx -1: random_short_side_scale_jitter, random_crop, horizontal_flip(0.5);
                                spatial_idx 0..5: scale jitter, uniform_crop(idx % 3), flip for 3..5
  random_short_side_scale_jitter (:35-79)   size = int(round(uniform(min, max))); short side -> size, long side ->
                                floor(long / short * size); F.interpolate(bilinear, align_corners=False); unchanged when
                                the short side already equals size
  random_crop (:101-134)        y = randint(0, H - size) if H > size, then x = randint(0, W - size) if W > size
  uniform_crop (:167-210)       centre offsets ceil((dim - size) / 2); idx 0 / 2 move along the LONGER side
  horizontal_flip (:137-164)    flip the last dim when uniform() < prob
Pinned against the reference itself: tests/golden/video_aug.npz is produced by tests/golden/gen_golden_video.py, which
imports /root/reference/datasets/video_transforms.py and records np.random's draws.
Colour jitter / grayscale (`--colorjitter`, `--use_grayscale`, default False, opt.py:47-50) are not on the default path.
"""
import math

import numpy as np
import torch

MEAN = [0.45, 0.45, 0.45]    # video_transforms.py:13-14
STD = [0.225, 0.225, 0.225]


def scaled_size(height, width, size):
    """(new_h, new_w) of random_short_side_scale_jitter for a drawn `size` (video_transforms.py:54-68)."""
    if (width <= height and width == size) or (height <= width and height == size):
        return height, width
    if width < height:
        return int(math.floor((float(height) / width) * size)), size
    return size, int(math.floor((float(width) / height) * size))


def draw_params(height, width, spatial_idx, min_scale, max_scale, crop_size, rng=np.random):
    """The reference's random draws in its order -> (new_h, new_w, y_off, x_off, flip)."""
    size = int(round(rng.uniform(min_scale, max_scale)))
    nh, nw = scaled_size(height, width, size)
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
            rng.uniform()   # horizontal_flip(1, .) still draws (video_transforms.py:157)
    return nh, nw, y, x, flip


def clip_augmentation_explicit(frames_u8, new_h, new_w, y_off, x_off, flip, crop_size):
    """frames_u8: uint8 [T,H,W,3] (torch or numpy) -> float32 [3,T,crop,crop] with the given parameters."""
    f = torch.as_tensor(np.asarray(frames_u8)).float()
    f = f / 255.0
    f = f - torch.tensor(MEAN)
    f = f / torch.tensor(STD)
    f = f.permute(0, 3, 1, 2).contiguous()
    if (new_h, new_w) != (f.shape[2], f.shape[3]):
        f = torch.nn.functional.interpolate(f, size=(new_h, new_w), mode="bilinear", align_corners=False)
    f = f[:, :, y_off:y_off + crop_size, x_off:x_off + crop_size]
    if flip:
        f = f.flip((-1))
    return f.permute(1, 0, 2, 3).contiguous()
