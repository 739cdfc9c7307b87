import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle.video_oracle import clip_augmentation_explicit, draw_params
from selavi_b200.video_transforms import clip_augmentation_batch
g = np.random.RandomState(7)
frames = g.randint(0, 256, size=(5, 8, 128, 171, 3)).astype(np.uint8)
params = [draw_params(128, 171, -1, 128, 160, 112, rng=g) for _ in range(5)]
refs = [clip_augmentation_explicit(frames[i], *p, 112) for i, p in enumerate(params)]
for mode in (0, 1, 2, 3):
    os.environ["SELAVI_CLIPAUG_FMA"] = str(mode)
    y = clip_augmentation_batch(torch.from_numpy(frames).cuda(), params=params, crop_size=112).cpu()
    errs = [float((y[i] - refs[i]).abs().max()) for i in range(5)]
    nz = [float(((y[i] - refs[i]) != 0).float().mean()) for i in range(5)]
    print("fma mode", mode, "max err", max(errs), "fraction of differing elements", max(nz))
