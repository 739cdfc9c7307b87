"""B200-native mirror of the reference `model.py` (AVModel / load_model; /root/reference/model.py:169-275).

The module tree, parameter names, shapes and the random-initialisation ORDER are those of the reference
(torchvision `r2plus1d_18` tv:video/resnet.py:45-121,184-300 + `ResNet(BasicBlock,[1,1,1,1])` tv:resnet.py:166-286
+ `MLPv2` model.py:62-90), so state_dicts are interchangeable and the same seed gives the same weights.  The
modules are parameter containers: `forward` never runs torch convolutions — the towers execute through
`engine.TowerRunner` (tcgen05 implicit-GEMM convs with fused BN/ReLU prologues, hand-written BN / pooling
kernels) and the heads through `engine.heads_forward`, all via the C ABI of include/selavi_b200.h.
"""
import torch
from torch import nn

from . import engine

__all__ = ["AVModel", "load_model", "get_model", "MLPv2", "VideoBaseNetwork", "AudioBaseNetwork"]


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.shape[0], -1)


class Unsqueeze(nn.Module):
    def forward(self, x):
        return x.unsqueeze(-1)


class Identity(nn.Module):
    def forward(self, x):
        return x


# ------------------------------------------------------------------------------------------------ video tower
class Conv2Plus1D(nn.Sequential):
    """(1x3x3 conv, BN, ReLU, 3x1x1 conv) — tv:video/resnet.py:45-61."""

    def __init__(self, in_planes, out_planes, midplanes, stride=1, padding=1):
        super().__init__(
            nn.Conv3d(in_planes, midplanes, kernel_size=(1, 3, 3), stride=(1, stride, stride),
                      padding=(0, padding, padding), bias=False),
            nn.BatchNorm3d(midplanes),
            nn.ReLU(inplace=True),
            nn.Conv3d(midplanes, out_planes, kernel_size=(3, 1, 1), stride=(stride, 1, 1), padding=(padding, 0, 0),
                      bias=False),
        )


class VideoBasicBlock(nn.Module):
    """tv:video/resnet.py:87-121."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        midplanes = (inplanes * planes * 3 * 3 * 3) // (inplanes * 3 * 3 + 3 * planes)
        super().__init__()
        self.conv1 = nn.Sequential(Conv2Plus1D(inplanes, planes, midplanes, stride), nn.BatchNorm3d(planes),
                                   nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(Conv2Plus1D(planes, planes, midplanes), nn.BatchNorm3d(planes))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class R2Plus1dStem(nn.Sequential):
    """tv:video/resnet.py:184-195."""

    def __init__(self):
        super().__init__(
            nn.Conv3d(3, 45, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False),
            nn.BatchNorm3d(45),
            nn.ReLU(inplace=True),
            nn.Conv3d(45, 64, kernel_size=(3, 1, 1), stride=(1, 1, 1), padding=(1, 0, 0), bias=False),
            nn.BatchNorm3d(64),
            nn.ReLU(inplace=True),
        )


class R2Plus1D18(nn.Module):
    """r2plus1d_18 with fc -> Identity (model.py:93-100), same construction / init order as torchvision."""

    def __init__(self):
        super().__init__()
        self.inplanes = 64
        self.stem = R2Plus1dStem()
        self.layer1 = self._make_layer(64, 2, stride=1)
        self.layer2 = self._make_layer(128, 2, stride=2)
        self.layer3 = self._make_layer(256, 2, stride=2)
        self.layer4 = self._make_layer(512, 2, stride=2)
        self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        fc = nn.Linear(512, 400)  # constructed (and initialised) by torchvision, then replaced by Identity
        for m in self.modules():  # tv:video/resnet.py:234-244
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(fc.weight, 0, 0.01)
        for m in self.modules():  # model.py:51-59 random_weight_init
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self.fc = Identity()

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes, kernel_size=1, stride=(stride, stride, stride), bias=False),
                nn.BatchNorm3d(planes))
        layers = [VideoBasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(VideoBasicBlock(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        return engine.tower_forward(self, "video", x)


# ------------------------------------------------------------------------------------------------ audio tower
class AudioBasicBlock(nn.Module):
    """tv:resnet.py:59-105."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class AudioResNet(nn.Module):
    """ResNet(BasicBlock, layers) with a 1-channel 7x7 conv1 and fc -> Identity (model.py:103-121)."""

    def __init__(self, layers=(1, 1, 1, 1)):
        super().__init__()
        self.inplanes = 64
        conv1_rgb = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)  # replaced below (model.py:117)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.layer4 = self._make_layer(512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        fc = nn.Linear(512, 1000)  # noqa: F841  (RNG parity with torchvision's constructor)
        nn.init.kaiming_normal_(conv1_rgb.weight, mode="fan_out", nonlinearity="relu")
        for m in self.modules():  # tv:resnet.py:209-214 (module order: bn1, layers...)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        conv1 = nn.Conv2d(1, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
        # keep torchvision's registration order of the state_dict: conv1 first
        mods = dict(self._modules)
        self._modules.clear()
        self._modules["conv1"] = conv1
        self._modules.update(mods)
        self.fc = Identity()

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [AudioBasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(AudioBasicBlock(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        return engine.tower_forward(self, "audio", x)


# ------------------------------------------------------------------------------------------------ heads
class MLPv2(nn.Module):
    """model.py:62-90.  `forward` runs the hand-written head kernels (single head)."""

    def __init__(self, n_input, n_classes, n_hidden=512, p=0.3):
        super().__init__()
        self.n_input, self.n_classes, self.n_hidden = n_input, n_classes, n_hidden
        if n_hidden is None:
            self.block_forward = nn.Sequential(Flatten(), nn.Dropout(p=p), nn.Linear(n_input, n_classes, bias=True))
        else:
            self.block_forward = nn.Sequential(
                Flatten(), nn.Dropout(p=p), nn.Linear(n_input, n_hidden, bias=False), Unsqueeze(),
                nn.BatchNorm1d(n_hidden), Flatten(), nn.ReLU(inplace=True), nn.Dropout(p=p),
                nn.Linear(n_hidden, n_classes, bias=True))

    def forward(self, x):
        return engine.heads_forward([self], x)[0]


class LinearHead(nn.Linear):
    """`use_mlp=False` head (model.py:207-208,217-219): an nn.Linear whose forward uses the batched head kernel."""

    def forward(self, x):
        return engine.heads_forward([self], x)[0]


def get_video_feature_extractor(vid_base_arch='r2plus1d_18', pretrained=False, duration=1):
    if vid_base_arch != 'r2plus1d_18':
        raise NotImplementedError("selavi_b200 implements the r2plus1d_18 video tower (the north-star path)")
    if pretrained:
        raise NotImplementedError("pretrained torchvision weights are not available offline; load a state_dict instead")
    print("Randomy initializing models")
    return R2Plus1D18()


def get_audio_feature_extractor(aud_base_arch='resnet18', pretrained=False, duration=1):
    assert aud_base_arch in ['resnet9', 'resnet18', 'resnet34', 'resnet50']
    if aud_base_arch == 'resnet9':
        print('resnet9, duration:', duration)
        return AudioResNet((1, 1, 1, 1))
    if aud_base_arch == 'resnet18':
        return AudioResNet((2, 2, 2, 2))
    if aud_base_arch == 'resnet34':
        return AudioResNet((3, 4, 6, 3))
    raise NotImplementedError("resnet50 (Bottleneck) audio tower is outside the hot path")


class VideoBaseNetwork(nn.Module):
    """model.py:135-149."""

    def __init__(self, vid_base_arch='r2plus1d_18', pretrained=False, norm_feat=False, duration=1):
        super().__init__()
        self.base = get_video_feature_extractor(vid_base_arch, pretrained=pretrained, duration=duration)
        self.norm_feat = norm_feat

    def forward(self, x):
        x = self.base(x).squeeze()
        if self.norm_feat:
            x = nn.functional.normalize(x, p=2, dim=1)
        return x


class AudioBaseNetwork(nn.Module):
    """model.py:152-166."""

    def __init__(self, aud_base_arch='resnet18', pretrained=False, norm_feat=False, duration=1):
        super().__init__()
        self.base = get_audio_feature_extractor(aud_base_arch, pretrained=pretrained, duration=duration)
        self.norm_feat = norm_feat

    def forward(self, x):
        x = self.base(x).squeeze()
        if self.norm_feat:
            x = nn.functional.normalize(x, p=2, dim=1)
        return x


class AVModel(nn.Module):
    """model.py:169-252 — same constructor, attributes and forward contract."""

    def __init__(self, vid_base_arch='r2plus1d_18', aud_base_arch='resnet9', pretrained=False, norm_feat=True,
                 use_mlp=False, headcount=1, num_classes=256, use_max_pool=False):
        super().__init__()
        self.use_mlp = use_mlp
        self.hc = headcount
        self.norm_feat = norm_feat
        self.return_features = False
        self.video_network = VideoBaseNetwork(vid_base_arch, pretrained=pretrained)
        self.audio_network = AudioBaseNetwork(aud_base_arch, pretrained=pretrained)
        encoder_dim = encoder_dim_a = n_hidden = 512
        if self.hc == 1:
            if use_mlp:
                print("Using MLP to be combined with SyncBN")
                self.mlp_v = MLPv2(encoder_dim, num_classes, n_hidden=n_hidden)
                self.mlp_a = MLPv2(encoder_dim_a, num_classes)
            else:
                print("Using Linear Layer")
                self.mlp_v = LinearHead(encoder_dim, num_classes)
                self.mlp_a = LinearHead(encoder_dim_a, num_classes)
        else:
            if use_mlp:
                print("Using MLP to be combined with SyncBN")
            for a in range(self.hc):
                if use_mlp:
                    setattr(self, "mlp_v%d" % a, MLPv2(encoder_dim, num_classes, n_hidden=n_hidden))
                    setattr(self, "mlp_a%d" % a, MLPv2(encoder_dim_a, num_classes))
                else:
                    setattr(self, "mlp_v%d" % a, LinearHead(encoder_dim, num_classes))
                    setattr(self, "mlp_a%d" % a, LinearHead(encoder_dim_a, num_classes))

    @property
    def _ddp_params_and_buffers_to_ignore(self):
        """Read by torch.nn.parallel.DistributedDataParallel when it is constructed around this model (main.py:156-160):
        the names it must neither broadcast nor reduce.  When the model was converted to SyncBatchNorm (main.py:117-118)
        these are every BatchNorm buffer (identical on all ranks by construction: the statistics are computed from the
        global batch) and the two towers' parameters, whose gradients the engine averages itself, overlapped with
        backward (engine.DDP_BYPASS).  DDP's construction-time broadcast of rank 0's state is done here for them."""
        import torch.distributed as dist
        if not (engine.DDP_BYPASS and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and any(isinstance(m, nn.SyncBatchNorm) for m in self.modules())):
            raise AttributeError("_ddp_params_and_buffers_to_ignore")
        towers = (("video_network.base", self.video_network.base, "video"), ("audio_network.base", self.audio_network.base, "audio"))
        names = [n for n, _ in self.named_buffers()]
        tensors = [b for _, b in self.named_buffers()]
        for prefix, net, kind in towers:
            names += [f"{prefix}.{n}" for n, _ in net.named_parameters()]
            tensors += [p.data for _, p in net.named_parameters()]
        # The side effects (switching the towers to engine-side gradient averaging, the state broadcast) happen only when
        # DistributedDataParallel's constructor is the reader: any other introspection (inspect.getmembers, dir loops)
        # just gets the names and cannot start a collective on one rank alone.
        import sys
        caller = sys._getframe(1)
        if caller.f_globals.get("__name__") != "torch.nn.parallel.distributed":
            return names
        for prefix, net, kind in towers:
            runner = net.__dict__.get("_sv_runner")
            if runner is None:
                runner = net.__dict__["_sv_runner"] = engine.TowerRunner(net, kind)
            runner.own_allreduce = True
        if tensors and not self.__dict__.get("_sv_ddp_synced"):
            # rank 0's state to every rank, like DDP's _sync_module_states (DDP reads this attribute twice: sync once)
            self.__dict__["_sv_ddp_synced"] = True
            # (dtypes in a FIXED order: a set's iteration order differs between processes, and a rank that broadcasts its
            # int64 counters while the others broadcast fp32 weights hangs the communicator — seen at 4 ranks)
            for dt in sorted({t.dtype for t in tensors}, key=str):
                grp = [t for t in tensors if t.dtype == dt]
                flat = torch.cat([t.reshape(-1) for t in grp])
                dist.broadcast(flat, 0)
                off = 0
                for t in grp:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()
        return names

    def _heads(self, prefix):
        if self.hc == 1:
            return [getattr(self, prefix)]
        return [getattr(self, "%s%d" % (prefix, h)) for h in range(self.hc)]

    def forward(self, img, spec, whichhead=0):
        if engine.AUDIO_STREAM and img.is_cuda and spec.is_cuda and img.device == spec.device:
            # the small audio tower overlaps the video tower on a second stream (its backward follows it there)
            main = torch.cuda.current_stream(img.device)
            side = engine.audio_stream(img.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                aud_features = self.audio_network(spec).squeeze()
            img_features = self.video_network(img).squeeze()
            main.wait_stream(side)
            spec.record_stream(side)
            aud_features.record_stream(main)
        else:
            img_features = self.video_network(img).squeeze()
            aud_features = self.audio_network(spec).squeeze()
        if self.return_features:
            return img_features, aud_features
        if len(aud_features.shape) == 1:
            aud_features = aud_features.unsqueeze(0)
        if len(img_features.shape) == 1:
            img_features = img_features.unsqueeze(0)
        outs1 = engine.heads_forward(self._heads("mlp_v"), img_features)   # all heads of a modality in one batch
        outs2 = engine.heads_forward(self._heads("mlp_a"), aud_features)
        if self.norm_feat:
            outs1 = [nn.functional.normalize(o, p=2, dim=1) for o in outs1]
            outs2 = [nn.functional.normalize(o, p=2, dim=1) for o in outs2]
        if self.hc == 1:
            return outs1[0], outs2[0]
        return outs1, outs2


def load_model(vid_base_arch='r2plus1d_18', aud_base_arch='resnet9', pretrained=False, norm_feat=True, use_mlp=False,
               headcount=1, num_classes=256, use_max_pool=False):
    """model.py:255-275."""
    return AVModel(vid_base_arch=vid_base_arch, aud_base_arch=aud_base_arch, pretrained=pretrained, norm_feat=norm_feat,
                   use_mlp=use_mlp, headcount=headcount, num_classes=num_classes, use_max_pool=use_max_pool)


get_model = load_model  # name used by BASELINE.json's north_star
