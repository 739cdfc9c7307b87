"""The reference's training-loop calling sequence (main.py:263-302: cluster() at the scheduled iterations, forward, get_loss on
the pseudo-labels, zero_grad, backward, optimizer.step) driven through the DROP-IN modules — `dropin/` ahead of everything on
sys.path, imported under the reference's own module names — on a small synthetic dataset.  /root/reference does not exist
on the GPU box, so main.py itself cannot be run there; this restates its loop around the same imports (main.py:25-39).
Run on a CUDA machine:  python tools/dropin_loop.py      (prints DROPIN_LOOP OK)"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dropin"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from model import load_model  # noqa: E402            (main.py:25)  -> dropin/model.py
from src.sk_utils import cluster  # noqa: E402        (main.py:27)  -> dropin/src/sk_utils.py
from selavi_b200.utils import get_loss  # noqa: E402  (main.py:29-39 imports it from utils)


class Clips(torch.utils.data.Dataset):
    """what AVideoDataset hands to the loop: (video, audio, label, index, video index) + the attributes cluster() reads"""

    def __init__(self, n, classes):
        r = np.random.default_rng(5)
        self.v = torch.from_numpy(r.standard_normal((n, 3, 4, 32, 32)).astype(np.float32) * np.linspace(0.5, 2, n, dtype=np.float32).reshape(n, 1, 1, 1, 1))
        self.a = torch.from_numpy((r.standard_normal((n, 1, 65, 40)) * 17.89 + 1.93).astype(np.float32))
        self._labels = r.integers(0, classes, n).tolist()
        self.valid_indices = np.arange(n)

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i], self.a[i], self._labels[i], i, i


class Log:
    def __init__(self):
        self.lines = []

    def info(self, msg, **_):
        self.lines.append(str(msg))


def main():
    torch.manual_seed(31)
    np.random.seed(31)
    hc, K, N = 2, 16, 96
    args = types.SimpleNamespace(world_size=1, rank=0, workers=0, ind_groups=1, headcount=hc, match=True, distribution="gauss",
                                 gauss_sd=0.1, diff_dist_per_head=True, diff_dist_every=False, dist=None, lamb=20, dump_path="")
    model = load_model(vid_base_arch="r2plus1d_18", aud_base_arch="resnet9", use_mlp=True, num_classes=K, pretrained=False,
                       norm_feat=False, use_max_pool=False, headcount=hc)                                  # main.py:105-114
    model = model.cuda()                                                                                    # main.py:126
    optimizer = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-5)               # main.py:132-137
    dataset = Clips(N, K)
    loader = torch.utils.data.DataLoader(dataset, batch_size=16, shuffle=True, drop_last=True)
    selflabels = torch.zeros((N, hc), dtype=torch.long, device="cuda")                                      # main.py:166
    logger, sk_counter, losses = Log(), 0, []
    model.train()
    it = 0
    for _epoch in range(2):
        for video, audio, _, selected, _ in loader:
            if it in (0, 7):                                                                               # main.py:276-281
                selflabels = cluster(args, selflabels, dataset, model, sk_counter, logger, None, None, it)
                sk_counter += 1
            video, audio = video.cuda(non_blocking=True), audio.cuda(non_blocking=True)                    # main.py:283-285
            feat_v, feat_a = model(video, audio)
            loss_vid = get_loss(feat_v, selflabels[selected, :], headcount=hc)                              # main.py:287-293
            loss_aud = get_loss(feat_a, selflabels[selected, :], headcount=hc)
            loss = 0.5 * loss_vid + 0.5 * loss_aud
            optimizer.zero_grad()                                                                           # main.py:296-302
            loss.backward()
            optimizer.step()
            losses.append(float(loss.item()))
            it += 1
    assert tuple(selflabels.shape) == (N, hc) and selflabels.dtype == torch.long and selflabels.is_cuda
    assert int(selflabels.min()) >= 0 and int(selflabels.max()) < K and selflabels[:, 0].unique().numel() > 1
    assert model.training and model.return_features is False
    assert all(np.isfinite(losses)) and len(losses) == 12, losses
    assert isinstance(args.dist, list) and len(args.dist) == hc     # Gaussian marginals drawn and kept (checkpointed by main.py:226)
    text = "\n".join(logger.lines)
    for needle in ("NMI_v:", "NMI-tolabels:", "aNMI-tolabels:", "Head 0, Cost", "Final Cost", "initial cost", "final cost"):
        assert needle in text, (needle, text[-800:])
    print(f"DROPIN_LOOP OK: loss {losses[0]:.4f} -> {losses[-1]:.4f}, {len(logger.lines)} log lines, SK calls {sk_counter}")


if __name__ == "__main__":
    main()
