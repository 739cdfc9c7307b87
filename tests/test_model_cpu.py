"""CPU checks of the model boundary: module tree / state_dict / initialisation parity with the reference, and
that the product path refuses to run without CUDA (no CPU fallback)."""
import pytest
import torch

from oracle import ref_loader
from selavi_b200 import model as sv_model


def test_state_dict_layout():
    m = sv_model.load_model(use_mlp=True, headcount=2, num_classes=28, norm_feat=False)
    sd = m.state_dict()
    assert len(sd) == 326
    for k in ["video_network.base.stem.0.weight", "video_network.base.stem.4.running_var",
              "video_network.base.layer2.0.conv1.0.3.weight", "video_network.base.layer3.0.downsample.1.num_batches_tracked",
              "audio_network.base.conv1.weight", "audio_network.base.layer4.0.downsample.0.weight",
              "mlp_v0.block_forward.2.weight", "mlp_a1.block_forward.8.bias", "mlp_v1.block_forward.4.running_mean"]:
        assert k in sd, k
    assert tuple(sd["video_network.base.layer2.0.conv1.0.0.weight"].shape) == (230, 64, 1, 3, 3)
    assert tuple(sd["video_network.base.layer4.1.conv2.0.3.weight"].shape) == (512, 1152, 3, 1, 1)
    assert tuple(sd["audio_network.base.conv1.weight"].shape) == (64, 1, 7, 7)
    assert sum(p.numel() for p in sv_model.load_model(use_mlp=True, headcount=10, num_classes=400).parameters()) == 45567005
    assert sv_model.get_model is sv_model.load_model
    # attributes poked by src/sk_utils.py / get_clusters.py
    assert m.return_features is False and m.hc == 2 and m.use_mlp is True
    assert isinstance(list(m.mlp_a0.modules())[-1], torch.nn.Linear)
    assert hasattr(m.video_network.base, "layer4") and hasattr(m.video_network.base, "stem")


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("kw", [dict(use_mlp=True, headcount=2, num_classes=28, norm_feat=False),
                                dict(use_mlp=False, headcount=1, num_classes=16),
                                dict(use_mlp=True, headcount=1, num_classes=309)])
def test_same_seed_same_weights_as_reference(kw):
    ref = ref_loader.load_model_module()
    torch.manual_seed(31)
    a = ref.load_model(**kw)
    torch.manual_seed(31)
    b = sv_model.load_model(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    # SyncBN conversion (main.py:117-118) keeps parameters and the tree usable
    c = torch.nn.SyncBatchNorm.convert_sync_batchnorm(b)
    assert list(c.state_dict().keys()) == list(sa.keys())


def test_no_cpu_fallback():
    m = sv_model.load_model(use_mlp=True, headcount=1, num_classes=8, norm_feat=False)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 2, 16, 16), torch.zeros(1, 1, 33, 20))
    from selavi_b200.utils import get_loss
    with pytest.raises(ValueError):
        get_loss(torch.zeros(2, 4), torch.zeros(2, dtype=torch.long))
