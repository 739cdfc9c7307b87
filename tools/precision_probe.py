"""Analysis-only probe (CPU): logit error of split-precision conv/linear operand schemes on the
reference-constructed model in train mode. Decides the MMA precision policy."""
import sys, copy, torch, torchvision
sys.path.insert(0, '/root/reference')
torchvision.models.resnet._resnet = lambda arch, block, layers, pretrained, progress, **kw: torchvision.models.resnet.ResNet(block, layers, **kw)
import model as refmodel
import torch.nn.functional as F

torch.manual_seed(31)
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 8
m = refmodel.load_model(norm_feat=False, use_mlp=True, headcount=1, num_classes=28)
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
m.train()
x = torch.randn(B, 3, T, 112, 112); s = torch.randn(B, 1, 257, 99) * 17.89 + 1.93
m64 = copy.deepcopy(m).double()
with torch.no_grad():
    ref64 = [o.clone() for o in m64(x.double(), s.double())]
    ref32 = [o.clone() for o in copy.deepcopy(m)(x, s)]

def trunc_tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
def rn_tf32(t):
    i = t.view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)
def split(t, mode):
    if mode.startswith('bf16'):
        hi = t.bfloat16().float(); lo = (t - hi).bfloat16().float()
        lo2 = (t - hi - lo).bfloat16().float()
        return hi, lo, lo2
    if mode.startswith('tf32t'):
        hi = trunc_tf32(t); lo = trunc_tf32(t - hi); return hi, lo, None
    hi = rn_tf32(t); lo = rn_tf32(t - hi); return hi, lo, None

def make_hook(mode):
    def hook(mod, inp):
        return None
    return hook

def patched_forward(mode):
    def fwd(self, x):
        w = self.weight
        def op(a, b):
            if isinstance(self, torch.nn.Linear): return F.linear(a, b)
            if isinstance(self, torch.nn.Conv3d): return F.conv3d(a, b, None, self.stride, self.padding)
            return F.conv2d(a, b, None, self.stride, self.padding)
        xh, xl, xl2 = split(x, mode); wh, wl, wl2 = split(w, mode)
        if mode.endswith('x1'): out = op(xh, wh)
        elif mode.endswith('x3'): out = op(xh, wh) + (op(xh, wl) + op(xl, wh))
        elif mode.endswith('x6'): out = op(xh, wh) + (op(xh, wl) + op(xl, wh)) + (op(xl, wl) + op(xh, wl2) + op(xl2, wh))
        if getattr(self, 'bias', None) is not None: out = out + self.bias
        return out
    return fwd

def rel(a, b): return ((a.double() - b).norm() / b.norm()).item()
print('fp32 ref vs fp64: v %.2e a %.2e' % (rel(ref32[0], ref64[0]), rel(ref32[1], ref64[1])))
for mode in ['bf16x1', 'tf32x1', 'bf16x3', 'bf16x6', 'tf32tx3', 'tf32x3']:
    mm = copy.deepcopy(m)
    for mod in mm.modules():
        if isinstance(mod, (torch.nn.Conv3d, torch.nn.Conv2d, torch.nn.Linear)):
            mod.forward = patched_forward(mode).__get__(mod)
    with torch.no_grad():
        o = mm(x, s)
    print('%s: vs fp64 v %.2e a %.2e | vs fp32ref v %.2e a %.2e | argmax agree v %.3f' % (
        mode, rel(o[0], ref64[0]), rel(o[1], ref64[1]), rel(o[0], ref32[0].double()), rel(o[1], ref32[1].double()),
        (o[0].argmax(1) == ref32[0].argmax(1)).float().mean().item()), flush=True)
