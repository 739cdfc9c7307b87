"""Execution engine of the B200-native SeLaVi hot path: runs the towers / heads / loss of `model.py` through
the C-ABI kernels and wires them into torch.autograd (so DDP's gradient hooks, torch optimizers and the
reference training loop main.py:263-302 work unchanged).

Data layout in HBM: activations channels-last fp32 [N,T,H,W,Cs] (Cs = channels padded to 8).  Every convolution
stores only its RAW output z; the following train-mode BatchNorm(+ReLU) lives as a per-channel (scale, shift)
pair that the NEXT convolution applies on the fly in its operand loader ("pending" activation), so normalised
activations are written to HBM only at residual joins (block outputs), after the stem and after the max-pool.
BN batch statistics come from the conv epilogue (per-tile partial sums), reduced in fp64.

No torch convolution / batch-norm / linear / cross-entropy call is made anywhere on this path, and there is no
CPU fallback: tensors must be CUDA and the shared library must be present.
"""
import os

import torch
import torch.distributed as dist
from torch import nn

from . import _lib, ops

# tf32x3 split (fp32-class accuracy) is the parity mode; SELAVI_MMA_PASSES=1 selects single-pass tf32 (fast mode).
PASSES = int(os.environ.get("SELAVI_MMA_PASSES", "3"))
# backward (data + weight gradients): "bf16x3" (default; gradient planes written as bf16 hi/lo by the BN-backward
# kernel, cp.async-fed tcgen05 kind::f16 kernels, measured 4e-6 per-layer error) or "tf32x3" (register-staged loaders)
BWD = os.environ.get("SELAVI_BWD", "bf16x3")
# forward of the convolutions the tap-reuse kernel does not cover (strided, 7x7, 1x1): "fp16x3" (default; fp16 hi/lo operands
# like conv_halo.cu, half the stages / shared-memory bytes / MMAs of tf32x3) or "tf32x3"
IGEMM_FWD = os.environ.get("SELAVI_IGEMM_FWD", "fp16x3")
# the bf16 hi/lo planes of a convolution's INPUT (operand of its weight gradient) are written by the BatchNorm-backward
# apply pass of the unit that produced that activation (it reads z anyway) instead of by a separate split pass that reads
# the fp32 tensor once more; the weight gradient is launched when they exist (one layer later in backward order)
PLANES_FROM_BN = os.environ.get("SELAVI_PLANES_FROM_BN", "1") == "1"
# the per-channel sums of a unit's BatchNorm-backward pass (sum g*mask, sum g*mask*zhat) are accumulated in the epilogue
# of the tap-reuse data-gradient kernel that PRODUCES g (it holds g in registers and reads the unit's z tile once) instead
# of by a separate pass that reads g and z again from HBM
FUSE_BN_BWD_STATS = os.environ.get("SELAVI_FUSE_BN_BWD_STATS", "1") == "1"


def _stream():
    return _lib.stream_ptr()


def _world(bn):
    """SyncBatchNorm semantics only when the module was converted (main.py:117-118) and a group is up."""
    if isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(bn.process_group) if bn.process_group is not None else dist.get_world_size()
    return 1


# weight gradients are off the critical path of backward (nothing downstream reads dW before the optimizer): launch them
# on a side stream so that they overlap the HBM-bound BatchNorm-backward passes and the data gradient of the layers below
WGRAD_STREAM = os.environ.get("SELAVI_WGRAD_STREAM", "1") == "1"
_side_streams = {}


def _side_stream(dev):
    """weight-gradient stream paired with the CURRENT stream (the two towers' backward passes run on different streams)"""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    s = _side_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev, priority=int(os.environ.get("SELAVI_WGRAD_PRIORITY", "-1")))
        _side_streams[key] = s
    return s


# The audio tower (1 % of the FLOPs, ~100 launch-latency-bound kernels per pass) runs on its own stream next to the
# video tower: forward here, backward by autograd on the same stream (backward nodes run on their forward stream).
AUDIO_STREAM = os.environ.get("SELAVI_AUDIO_STREAM", "1") == "1"
_audio_streams = {}


def audio_stream(dev):
    s = _audio_streams.get(dev.index)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _audio_streams[dev.index] = s
    return s


# DistributedDataParallel (main.py:156-160) is told to ignore the towers' parameters and every BatchNorm buffer
# (model.AVModel._ddp_params_and_buffers_to_ignore); the engine averages the tower gradients itself: chunks of ~32 MB are
# flattened and all-reduced (NCCL, AVG) on a communication stream as soon as the weight gradients of the blocks they
# belong to are complete, so the exchange of layer 4 (3/4 of all parameters, finished in the first tenth of backward)
# hides behind layers 3..1.  DDP's own path hands autograd's 247 gradients to the reducer only when the tower's single
# autograd node returns (nothing overlaps), copies each of them into a bucket and back, and re-broadcasts 300 BatchNorm
# buffers before every forward although SyncBatchNorm keeps them identical by construction.  SELAVI_DDP_BYPASS=0 restores it.
DDP_BYPASS = os.environ.get("SELAVI_DDP_BYPASS", "1") == "1"
REDUCE_CHUNK_BYTES = int(os.environ.get("SELAVI_REDUCE_CHUNK_MB", "32")) << 20
_comm_streams = {}


def _comm_stream(dev):
    s = _comm_streams.get(dev.index)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _comm_streams[dev.index] = s
    return s


# cross-rank exchange of the SyncBN statistic vectors: "p2p" = one tiny kernel over NVSwitch peer memory (default inside
# one node, world <= 8), "nccl" = torch.distributed.all_reduce
BN_EXCHANGE = os.environ.get("SELAVI_BN_EXCHANGE", "p2p")
_p2p = {}


def _allreduce(t, bn):
    group = bn.process_group
    if BN_EXCHANGE == "p2p":
        key = id(group)
        ar = _p2p.get(key)
        if ar is None:
            from .symm import P2PAllReduce
            try:
                ar = P2PAllReduce(group)
            except _lib.SelaviError as e:
                # P2PAllReduce / SymmetricBuffer agree on availability collectively: this branch is taken by EVERY rank
                # of the group or by none, so the NCCL fallback below cannot be mixed with spinning P2P kernels
                import warnings
                warnings.warn(f"selavi_b200: NVSwitch P2P exchange of the SyncBatchNorm statistics unavailable ({e}); "
                              "falling back to torch.distributed.all_reduce (NCCL)")
                ar = False
            _p2p[key] = ar
        if ar:
            ar.allreduce_(t.view(-1))
            return
    dist.all_reduce(t, group=group)


class Act:
    """An activation: a materialised tensor, or a raw conv output with a pending per-channel affine (+ReLU).
    `key`: data pointer of the tensor through which the BatchNorm-backward pass of the PRODUCING unit will see this
    activation (its own z for a pending activation / the stem output, the materialised block output otherwise); None for
    tensors no BatchNorm unit produces (network input, max-pool output).  Weight gradients whose input has a key are
    deferred until that pass has emitted the activation's bf16 hi/lo planes (PLANES_FROM_BN)."""
    __slots__ = ("t", "scale", "shift", "relu", "c", "key", "unit")

    def __init__(self, t, c, scale=None, shift=None, relu=False, key=None, unit=None):
        self.t, self.c, self.scale, self.shift, self.relu, self.key = t, c, scale, shift, relu, key
        self.unit = unit      # ConvRec of the conv+BN unit whose PENDING activation this is (its backward consumes our dgrad)


class ConvRec:
    """Everything the backward pass needs about one conv+BN unit."""
    __slots__ = ("conv", "bn", "geom", "inp", "z", "scale", "shift", "mean", "invstd", "count", "fused_stats")


def _packed_dgrad_bf16(conv, geom):
    weight = conv.weight
    cache = conv.__dict__.setdefault("_sv_pack", {})
    tag = (weight.data_ptr(), weight._version)
    hit = cache.get("dgrad_bf16")
    if hit is not None and hit[0] == tag:
        return hit[1]
    buf = ops.pack_weights_dgrad_bf16(weight, geom, out=hit[1] if (hit is not None and hit[1].device == weight.device) else None)
    cache["dgrad_bf16"] = (tag, buf)
    return buf


def _packed(conv, geom, mode):
    """Packed (pre-tiled, pre-swizzled, tf32 hi/lo) B operand of `conv.weight`, cached ON the module and
    re-packed whenever the parameter storage or its version counter changes (optimizer step, load_state_dict,
    .to(device), match_order's row permutation)."""
    weight = conv.weight
    cache = conv.__dict__.setdefault("_sv_pack", {})
    tag = (weight.data_ptr(), weight._version)
    hit = cache.get(mode)
    if hit is not None and hit[0] == tag:
        return hit[1]
    prev = hit[1] if (hit is not None and hit[1].device == weight.device) else None
    if isinstance(mode, tuple):
        buf = ops.pack_weights_halo(weight, geom, out=prev, mode=1 if mode[0] == "halo_dgrad" else 0)
    else:
        buf = ops.pack_weights(weight, geom, mode, out=prev)
    cache[mode] = (tag, buf)
    return buf


def _geom_of(conv, nb, thw):
    if isinstance(conv, nn.Conv3d):
        k, s, p = conv.kernel_size, conv.stride, conv.padding
    else:
        k, s, p = (1,) + tuple(conv.kernel_size), (1,) + tuple(conv.stride), (0,) + tuple(conv.padding)
    return ops.ConvGeom(nb, conv.in_channels, conv.out_channels, thw, k, s, p)


class TowerRunner:
    """Forward / backward of one encoder tower (video R(2+1)D-18 or audio ResNet) on the C-ABI kernels."""

    def __init__(self, net, kind):
        self.kind = kind
        self.own_allreduce = None      # process group (or True for the default group) once DDP ignores this tower's parameters

    # ------------------------------------------------------------------ forward pieces
    def conv_bn(self, act, conv, bn, training, tape):
        lib = _lib.lib()
        x = act.t
        nb, t, h, w, _ = x.shape
        geom = _geom_of(conv, nb, (t, h, w))
        plan = ops.halo_plan(geom) if PASSES == 3 else None
        halo = plan is not None
        f16 = PASSES == 3 and IGEMM_FWD == "fp16x3"
        wp = _packed(conv, geom, ("halo", plan[1], plan[2]) if halo else (2 if f16 else 0))
        igemm_passes = 6 if f16 else PASSES
        dev = x.device
        cs = geom.cos
        scale = torch.empty(cs, dtype=torch.float32, device=dev)
        shift = torch.empty(cs, dtype=torch.float32, device=dev)
        rec = None
        if training:
            stats = ops.stats_buffer(geom, dev, halo=halo)
            if halo:
                z = ops.conv_forward_halo(x, wp, geom, scale=act.scale, shift=act.shift, relu=act.relu, stats=stats)
            else:
                z = ops.conv_forward(x, wp, geom, scale=act.scale, shift=act.shift, relu=act.relu, stats=stats, passes=igemm_passes)
            sums = torch.empty(2 * cs, dtype=torch.float64, device=dev)
            _lib.check(lib.selavi_bn_reduce_partials(_lib.ptr(stats), stats.shape[0], stats.shape[2], cs, _lib.ptr(sums),
                                                     _stream()), "selavi_bn_reduce_partials")
            count = float(geom.m_out)
            world = _world(bn)
            if world > 1:
                _allreduce(sums, bn)
                count *= world
            mean = torch.empty(cs, dtype=torch.float32, device=dev)
            invstd = torch.empty(cs, dtype=torch.float32, device=dev)
            track = bn.track_running_stats and bn.running_mean is not None
            mom = 0.1 if bn.momentum is None else bn.momentum
            _lib.check(lib.selavi_bn_finalize(_lib.ptr(sums), count, _lib.ptr(bn.weight), _lib.ptr(bn.bias),
                                              _lib.ptr(bn.running_mean) if track else None,
                                              _lib.ptr(bn.running_var) if track else None, mom, bn.eps, geom.co, cs,
                                              _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(mean), _lib.ptr(invstd),
                                              1 if track else 0, _stream()), "selavi_bn_finalize")
            if track:
                self._nbt.append(bn.num_batches_tracked)
            if tape is not None:
                rec = ConvRec()
                rec.conv, rec.bn, rec.geom, rec.inp, rec.z = conv, bn, geom, act, z
                rec.scale, rec.shift, rec.mean, rec.invstd, rec.count = scale, shift, mean, invstd, count
                rec.fused_stats = None
        else:
            if halo:
                z = ops.conv_forward_halo(x, wp, geom, scale=act.scale, shift=act.shift, relu=act.relu, stats=None)
            else:
                z = ops.conv_forward(x, wp, geom, scale=act.scale, shift=act.shift, relu=act.relu, stats=None, passes=igemm_passes)
            _lib.check(lib.selavi_bn_eval_affine(_lib.ptr(bn.weight), _lib.ptr(bn.bias), _lib.ptr(bn.running_mean),
                                                 _lib.ptr(bn.running_var), bn.eps, geom.co, cs, _lib.ptr(scale),
                                                 _lib.ptr(shift), _stream()), "selavi_bn_eval_affine")
        return z, scale, shift, rec, geom

    @staticmethod
    def bn_apply(z, scale, shift, res=None, rscale=None, rshift=None, relu=True):
        out = torch.empty_like(z)
        m = z.numel() // z.shape[-1]
        _lib.check(_lib.lib().selavi_bn_apply(_lib.ptr(z), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(res), _lib.ptr(rscale),
                                              _lib.ptr(rshift), 1 if relu else 0, _lib.ptr(out), m, z.shape[-1], _stream()),
                   "selavi_bn_apply")
        return out

    def block(self, x_act, main, downsample, training, tape):
        """main: list of (conv, bn) of the residual branch; downsample: (conv, bn) or None.  Returns block output."""
        act = x_act
        recs = []
        z = scale = shift = None
        for i, (conv, bn) in enumerate(main):
            z, scale, shift, rec, geom = self.conv_bn(act, conv, bn, training, tape)
            recs.append(rec)
            act = Act(z, geom.co, scale, shift, relu=True, key=z.data_ptr(), unit=rec)
        rd = None
        if downsample is not None:
            zd, sd, bd, rd, _ = self.conv_bn(x_act, downsample[0], downsample[1], training, tape)
            y = self.bn_apply(z, scale, shift, res=zd, rscale=sd, rshift=bd, relu=True)
        else:
            y = self.bn_apply(z, scale, shift, res=x_act.t, relu=True)
        if tape is not None:
            tape.append(("block", recs, rd, x_act, y))
        return Act(y, act.c, key=y.data_ptr())

    def forward(self, net, x, training, tape):
        lib = _lib.lib()
        self._nbt = []
        if not (x.is_cuda and x.dtype == torch.float32):
            raise ValueError("selavi_b200 towers need float32 CUDA input (no CPU fallback)")
        x = x.contiguous()
        if self.kind == "video":
            nb, c, t, h, w = x.shape
        else:
            nb, c, h, w = x.shape
            t = 1
        cs = ops.padc(c)
        x_cl = torch.empty((nb, t, h, w, cs), dtype=torch.float32, device=x.device)
        _lib.check(lib.selavi_nchw_to_cl(_lib.ptr(x), _lib.ptr(x_cl), nb, c, t * h * w, cs, _stream()), "selavi_nchw_to_cl")
        act = Act(x_cl, c)
        if self.kind == "video":
            stem = net.stem
            z0, s0, b0, r0, g0 = self.conv_bn(act, stem[0], stem[1], training, tape)
            z1, s1, b1, r1, g1 = self.conv_bn(Act(z0, g0.co, s0, b0, True, key=z0.data_ptr(), unit=r0), stem[3], stem[4], training, tape)
            a = self.bn_apply(z1, s1, b1, relu=True)
            if tape is not None:
                tape.append(("vstem", r0, r1))
            act = Act(a, g1.co, key=z1.data_ptr())     # a == relu(bn(z1)): the stem unit r1 re-creates it in backward
            for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
                for blk in layer:
                    main = [(blk.conv1[0][0], blk.conv1[0][1]), (blk.conv1[0][3], blk.conv1[1]),
                            (blk.conv2[0][0], blk.conv2[0][1]), (blk.conv2[0][3], blk.conv2[1])]
                    ds = (blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                    act = self.block(act, main, ds, training, tape)
        else:
            z0, s0, b0, r0, g0 = self.conv_bn(act, net.conv1, net.bn1, training, tape)
            ho, wo = (g0.ho + 2 - 3) // 2 + 1, (g0.wo + 2 - 3) // 2 + 1
            pooled = torch.empty((nb, 1, ho, wo, g0.cos), dtype=torch.float32, device=x.device)
            _lib.check(lib.selavi_maxpool3x3s2_fwd(_lib.ptr(z0), _lib.ptr(s0), _lib.ptr(b0), _lib.ptr(pooled), nb, g0.ho,
                                                   g0.wo, g0.cos, _stream()), "selavi_maxpool3x3s2_fwd")
            if tape is not None:
                tape.append(("astem", r0, s0, b0))
            act = Act(pooled, g0.co)
            for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
                for blk in layer:
                    main = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
                    ds = (blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                    act = self.block(act, main, ds, training, tape)
        y = act.t
        p = y.shape[1] * y.shape[2] * y.shape[3]
        feat = torch.empty((nb, act.c), dtype=torch.float32, device=x.device)
        _lib.check(lib.selavi_avgpool_fwd(_lib.ptr(y), _lib.ptr(feat), nb, p, y.shape[-1], act.c, _stream()), "selavi_avgpool_fwd")
        if tape is not None:
            tape.append(("pool", tuple(y.shape), act.c))
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)
        return feat

    # ------------------------------------------------------------------ backward pieces
    def conv_bn_bwd(self, rec, g, mask_mode, grads, act_mask=None, want_dx=True, dx_out=None, dx_accumulate=False,
                    gres=None, gres_accumulate=False):
        """g: gradient wrt the (activated) output of this conv+BN unit.  Returns gradient wrt its input activation."""
        lib = _lib.lib()
        geom, z = rec.geom, rec.z
        dev = z.device
        cs, M = geom.cos, geom.m_out
        sums = torch.empty(2 * cs, dtype=torch.float64, device=dev)
        fused, rec.fused_stats = rec.fused_stats, None
        if fused is not None and mask_mode == 2:
            # the data-gradient kernel that wrote g left the per-tile partial sums of this very reduction
            _lib.check(lib.selavi_bn_reduce_partials(_lib.ptr(fused), fused.shape[0], fused.shape[2], cs, _lib.ptr(sums),
                                                     _stream()), "selavi_bn_reduce_partials")
        else:
            nblk = lib.selavi_bn_bwd_blocks(M)
            partial = torch.empty(nblk * 2 * cs, dtype=torch.float32, device=dev)
            _lib.check(lib.selavi_bn_bwd_reduce(_lib.ptr(g), _lib.ptr(z), _lib.ptr(act_mask), mask_mode, _lib.ptr(rec.scale),
                                                _lib.ptr(rec.shift), _lib.ptr(rec.mean), _lib.ptr(rec.invstd), M, cs,
                                                _lib.ptr(partial), _lib.ptr(sums), _stream()), "selavi_bn_bwd_reduce")
        bn = rec.bn
        if bn.weight is not None and bn.weight.requires_grad:
            s32 = sums.view(2, cs)[:, :geom.co].float()
            grads[bn.bias] = s32[0].contiguous()
            grads[bn.weight] = s32[1].contiguous()
        if _world(bn) > 1:
            sums = sums.clone()
            _allreduce(sums, bn)
        bf16 = BWD == "bf16x3"
        conv, inp = rec.conv, rec.inp
        need_dw = conv.weight.requires_grad
        dz = z_hi = z_lo = None
        if bf16:
            z_hi = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
            z_lo = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
        else:
            dz = torch.empty_like(z)
        # weight gradients of the convolutions that consumed THIS unit's activation are waiting for its bf16 planes
        my_key = (act_mask.data_ptr() if mask_mode == 1 else z.data_ptr()) if mask_mode in (1, 2) else None
        waiting = self._pending.pop(my_key, None) if my_key is not None else None
        a_hi = a_lo = None
        if waiting:
            a_hi = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
            a_lo = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.selavi_bn_bwd_apply(_lib.ptr(g), _lib.ptr(z), _lib.ptr(act_mask), mask_mode, _lib.ptr(rec.scale),
                                           _lib.ptr(rec.shift), _lib.ptr(rec.mean), _lib.ptr(rec.invstd), _lib.ptr(sums),
                                           rec.count, M, cs, _lib.ptr(dz), _lib.ptr(gres), 1 if gres_accumulate else 0,
                                           _lib.ptr(z_hi), _lib.ptr(z_lo), _lib.ptr(a_hi), _lib.ptr(a_lo), _stream()), "selavi_bn_bwd_apply")
        for job in waiting or ():
            self._launch_wgrad(job, grads, a_hi, a_lo)
        if need_dw:
            job = (conv, geom, inp, z_hi, z_lo, dz)
            if bf16 and PLANES_FROM_BN and inp.key is not None:
                self._pending.setdefault(inp.key, []).append(job)
            else:
                self._launch_wgrad(job, grads, None, None)
        if not want_dx:
            return None
        if bf16:
            plan = ops.halo_plan(geom, 1) if PASSES == 3 else None
            if plan is not None:
                wpd = _packed(conv, geom, ("halo_dgrad", plan[1], plan[2]))
                prev = inp.unit
                # measured (tools/dgrad_fused_bench.py, profiles/r02_dgrad_fused_stats.txt): pays on the spatial convs (narrow
                # output, long MMA phase per tile: +0.02 ms against a 0.16 ms reduce pass on layer 1) and LOSES on the temporal
                # ones (144..1152-wide output, epilogue-bound: +0.47 ms against 0.32 ms), so only kt == 1 is fused
                if (FUSE_BN_BWD_STATS and geom.kt == 1 and prev is not None and dx_out is None and not dx_accumulate
                        and prev.z is inp.t):
                    dx, prev.fused_stats = ops.conv_dgrad_halo(z_hi, z_lo, wpd, geom,
                                                               bn=(prev.z, prev.scale, prev.shift, prev.mean, prev.invstd))
                    return dx
                return ops.conv_dgrad_halo(z_hi, z_lo, wpd, geom, out=dx_out, accumulate=dx_accumulate)
            return ops.conv_dgrad_bf16(z_hi, z_lo, _packed_dgrad_bf16(conv, geom), geom, out=dx_out, accumulate=dx_accumulate,
                                       passes=3 if PASSES == 3 else 1)
        wpt = _packed(conv, geom, 1)
        return ops.conv_dgrad(dz, wpt, geom, out=dx_out, accumulate=dx_accumulate, passes=PASSES)

    def _launch_wgrad(self, job, grads, a_hi, a_lo):
        """Weight gradient of one convolution, on the side stream paired with the current stream.  a_hi/a_lo: bf16 planes
        of the conv input (from the producing unit's BatchNorm-backward pass) or None (split inside the call)."""
        conv, geom, inp, z_hi, z_lo, dz = job
        dev = conv.weight.device
        dw = torch.empty_like(conv.weight)
        passes = 3 if PASSES == 3 else 1

        def run():
            if a_hi is not None:
                ops.conv_wgrad_bf16_planes(a_hi, a_lo, z_hi, z_lo, geom, dw, passes=passes)
            else:
                ops.conv_wgrad_bf16(inp.t, z_hi, z_lo, geom, dw, scale=inp.scale, shift=inp.shift, relu=inp.relu, passes=passes)

        if z_hi is None:     # SELAVI_BWD=tf32x3
            ops.conv_wgrad(inp.t, dz, geom, dw, scale=inp.scale, shift=inp.shift, relu=inp.relu, passes=13 if PASSES == 3 else 11)
        elif WGRAD_STREAM:
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                run()
            for t in (inp.t, z_hi, z_lo, dw, inp.scale, inp.shift, a_hi, a_lo):
                if t is not None:
                    t.record_stream(side)
            self._side_busy = True
        else:
            run()
        grads[conv.weight] = dw

    def backward(self, tape, dfeat, grads):
        self._side_busy = False
        self._comm_busy = False
        self._reduced = set()
        self._pending = {}
        try:
            self._backward(tape, dfeat, grads)
            for jobs in list(self._pending.values()):      # (no producing unit came by: cannot happen for these towers)
                for job in jobs:
                    self._launch_wgrad(job, grads, None, None)
            self._pending = {}
            self._reduce_ready(grads, dfeat.device, final=True)
        finally:
            main = torch.cuda.current_stream(dfeat.device)
            if self._side_busy:   # the weight gradients must be complete before autograd hands them to DDP / the optimizer
                main.wait_stream(_side_stream(dfeat.device))
            if self._comm_busy:
                main.wait_stream(_comm_stream(dfeat.device))

    def _reduce_ready(self, grads, dev, final=False):
        """Average the gradients produced since the last call over the data-parallel ranks (see DDP_BYPASS)."""
        if self.own_allreduce is None or not (dist.is_available() and dist.is_initialized()):
            return
        group = None if self.own_allreduce is True else self.own_allreduce
        if dist.get_world_size(group) < 2:
            return
        new = [t for p, t in grads.items() if id(p) not in self._reduced and t is not None]
        if not new or (not final and 4 * sum(t.numel() for t in new) < REDUCE_CHUNK_BYTES):
            return
        self._reduced.update(id(p) for p in grads)
        main, comm = torch.cuda.current_stream(dev), _comm_stream(dev)
        comm.wait_stream(main)                       # BatchNorm scale / bias gradients are produced on this stream
        if self._side_busy:
            comm.wait_stream(_side_stream(dev))      # ... the weight gradients on the side stream
        with torch.cuda.stream(comm):
            flat = torch.cat([t.reshape(-1) for t in new])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            off, pieces = 0, []
            for t in new:
                pieces.append(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
            torch._foreach_copy_(new, pieces)
        for t in new:
            t.record_stream(comm)
        self._comm_busy = True

    def _backward(self, tape, dfeat, grads):
        lib = _lib.lib()
        dfeat = dfeat.contiguous()
        g = None
        for entry in reversed(tape):
            self._reduce_ready(grads, dfeat.device)      # whatever the previous entries produced, once a chunk is full
            kind = entry[0]
            if kind == "pool":
                shape, c = entry[1], entry[2]
                g = torch.empty(shape, dtype=torch.float32, device=dfeat.device)
                p = shape[1] * shape[2] * shape[3]
                _lib.check(lib.selavi_avgpool_bwd(_lib.ptr(dfeat), _lib.ptr(g), shape[0], p, shape[-1], c, _stream()),
                           "selavi_avgpool_bwd")
            elif kind == "block":
                _, recs, rd, x_act, y = entry
                dx = torch.empty_like(x_act.t)
                if rd is not None:
                    self.conv_bn_bwd(rd, g, 1, grads, act_mask=y, dx_out=dx, dx_accumulate=False)
                    d = self.conv_bn_bwd(recs[-1], g, 1, grads, act_mask=y)
                else:
                    d = self.conv_bn_bwd(recs[-1], g, 1, grads, act_mask=y, gres=dx, gres_accumulate=False)
                for rec in reversed(recs[1:-1]):
                    d = self.conv_bn_bwd(rec, d, 2, grads)
                self.conv_bn_bwd(recs[0], d, 2, grads, dx_out=dx, dx_accumulate=True)
                g = dx
            elif kind == "vstem":
                _, r0, r1 = entry
                d = self.conv_bn_bwd(r1, g, 2, grads)
                self.conv_bn_bwd(r0, d, 2, grads, want_dx=False)
                g = None
            elif kind == "astem":
                _, r0, s0, b0 = entry
                z0 = r0.z
                da = torch.empty_like(z0)
                _lib.check(lib.selavi_maxpool3x3s2_bwd(_lib.ptr(g), _lib.ptr(z0), _lib.ptr(s0), _lib.ptr(b0), _lib.ptr(da),
                                                       z0.shape[0], z0.shape[2], z0.shape[3], z0.shape[4], _stream()),
                           "selavi_maxpool3x3s2_bwd")
                self.conv_bn_bwd(r0, da, 2, grads, want_dx=False)
                g = None


def _tower_params(net):
    return [p for p in net.parameters()]


class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, net, x, *params):
        tape = []
        feat = runner.forward(net, x, True, tape)
        ctx.runner, ctx.tape, ctx.params = runner, tape, params
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        grads = {}
        ctx.runner.backward(ctx.tape, dfeat, grads)
        ctx.tape = None
        return (None, None, None) + tuple(grads.get(p) for p in ctx.params)


def tower_forward(net, kind, x):
    """Forward of `R2Plus1D18` / `AudioResNet` (replaces tv:video/resnet.py:246-260 / tv:resnet.py:266-284)."""
    runner = net.__dict__.get("_sv_runner")
    if runner is None:
        runner = TowerRunner(net, kind)
        net.__dict__["_sv_runner"] = runner
    params = _tower_params(net)
    training = net.training
    if training and torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _TowerFn.apply(runner, net, x, *params)
    if torch.is_grad_enabled() and not training and any(p.requires_grad for p in params) and x.requires_grad:
        raise NotImplementedError("gradients through an eval-mode tower are outside the hot path")
    with torch.no_grad():
        return runner.forward(net, x, training, None)


# ====================================================================================================== heads
_tbl_cache = {}


def _ptr_table(tensors):
    key = tuple(t.data_ptr() for t in tensors)
    tbl = _tbl_cache.get(key)
    if tbl is None:
        tbl = torch.tensor(list(key), dtype=torch.int64, device=tensors[0].device)
        if len(_tbl_cache) > 4096:
            _tbl_cache.clear()
        _tbl_cache[key] = tbl
    return tbl


def _head_parts(head):
    """-> (lin1 or None, bn or None, lin2, p_drop)"""
    if isinstance(head, nn.Linear):
        return None, None, head, 0.0
    seq = head.block_forward
    if head.n_hidden is None:
        return None, None, seq[2], seq[1].p
    return seq[2], seq[4], seq[8], seq[1].p


def _bgemm(H, M, N, K, A=None, A_tbl=None, a=(0, 0, 0), Amask=None, am_bs=0, B=None, B_tbl=None, b=(0, 0, 0), bias=None,
           bias_tbl=None, bias_bs=0, C=None, c=(0, 0, 0), accumulate=False):
    _lib.check(_lib.lib().selavi_bgemm(H, M, N, K, _lib.ptr(A), _lib.ptr(A_tbl), a[0], a[1], a[2], _lib.ptr(Amask), am_bs,
                                       _lib.ptr(B), _lib.ptr(B_tbl), b[0], b[1], b[2], _lib.ptr(bias), _lib.ptr(bias_tbl),
                                       bias_bs, _lib.ptr(C), c[0], c[1], c[2], 1 if accumulate else 0, _stream()),
               "selavi_bgemm")


class _HeadsRun:
    """Batched forward/backward of H structurally identical heads sharing one input [B, F]."""

    def __init__(self, heads, training):
        self.heads = heads
        self.H = len(heads)
        parts = [_head_parts(h) for h in heads]
        self.lin1 = [p[0] for p in parts]
        self.bn = [p[1] for p in parts]
        self.lin2 = [p[2] for p in parts]
        self.p = parts[0][3]
        self.mlp = self.lin1[0] is not None
        self.training = training
        self.params = []
        for l1, bn, l2 in zip(self.lin1, self.bn, self.lin2):
            if l1 is not None:
                self.params += [l1.weight, bn.weight, bn.bias]
            self.params += [l2.weight] + ([l2.bias] if l2.bias is not None else [])

    def _mask(self, shape, dev):
        if not self.training or self.p <= 0.0:
            return None
        keep = 1.0 - self.p
        return torch.empty(shape, dtype=torch.float32, device=dev).bernoulli_(keep).div_(keep)

    def forward(self, x, save):
        lib = _lib.lib()
        x = x.contiguous()
        H, (B, F) = self.H, x.shape
        dev = x.device
        K = self.lin2[0].out_features
        w2 = _ptr_table([l.weight for l in self.lin2])
        b2 = _ptr_table([l.bias for l in self.lin2]) if self.lin2[0].bias is not None else None
        logits = torch.empty((H, B, K), dtype=torch.float32, device=dev)
        m1 = self._mask((H, B, F), dev)
        st = {"x": x, "m1": m1}
        if self.mlp:
            Fh = self.lin1[0].out_features
            w1 = _ptr_table([l.weight for l in self.lin1])
            z1 = torch.empty((H, B, Fh), dtype=torch.float32, device=dev)
            # z1[h] = (x * m1[h]) @ W1[h]^T
            _bgemm(H, B, Fh, F, A=x, a=(0, F, 1), Amask=m1, am_bs=B * F, B_tbl=w1, b=(0, 1, F), C=z1, c=(B * Fh, Fh, 1))
            scale = torch.empty((H, Fh), dtype=torch.float32, device=dev)
            shift = torch.empty_like(scale)
            gam, bet = _ptr_table([b.weight for b in self.bn]), _ptr_table([b.bias for b in self.bn])
            rm, rv = _ptr_table([b.running_mean for b in self.bn]), _ptr_table([b.running_var for b in self.bn])
            bn0 = self.bn[0]
            if self.training:
                sums = torch.empty((H, 2, Fh), dtype=torch.float64, device=dev)
                _lib.check(lib.selavi_heads_bn_stats(_lib.ptr(z1), H, B, Fh, _lib.ptr(sums), _stream()), "selavi_heads_bn_stats")
                count = float(B)
                world = _world(bn0)
                if world > 1:
                    _allreduce(sums, bn0)
                    count *= world
                mean = torch.empty_like(scale)
                invstd = torch.empty_like(scale)
                mom = 0.1 if bn0.momentum is None else bn0.momentum
                _lib.check(lib.selavi_heads_bn_finalize(_lib.ptr(sums), count, _lib.ptr(gam), _lib.ptr(bet), _lib.ptr(rm),
                                                        _lib.ptr(rv), mom, bn0.eps, H, Fh, _lib.ptr(scale), _lib.ptr(shift),
                                                        _lib.ptr(mean), _lib.ptr(invstd), 1, _stream()),
                           "selavi_heads_bn_finalize")
                torch._foreach_add_([b.num_batches_tracked for b in self.bn], 1)
                st.update(mean=mean, invstd=invstd, count=count)
            else:
                _lib.check(lib.selavi_heads_bn_eval_affine(_lib.ptr(gam), _lib.ptr(bet), _lib.ptr(rm), _lib.ptr(rv), bn0.eps,
                                                           H, Fh, _lib.ptr(scale), _lib.ptr(shift), _stream()),
                           "selavi_heads_bn_eval_affine")
            m2 = self._mask((H, B, Fh), dev)
            a1 = torch.empty_like(z1)
            _lib.check(lib.selavi_heads_act(_lib.ptr(z1), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(m2), _lib.ptr(a1), H, B,
                                            Fh, _stream()), "selavi_heads_act")
            _bgemm(H, B, K, Fh, A=a1, a=(B * Fh, Fh, 1), B_tbl=w2, b=(0, 1, Fh), bias_tbl=b2, C=logits, c=(B * K, K, 1))
            st.update(z1=z1, a1=a1, m2=m2, scale=scale, shift=shift, w1=w1)
        else:
            _bgemm(H, B, K, F, A=x, a=(0, F, 1), Amask=m1, am_bs=B * F, B_tbl=w2, b=(0, 1, F), bias_tbl=b2, C=logits,
                   c=(B * K, K, 1))
        st["w2"] = w2
        return logits, (st if save else None)

    def backward(self, st, dlogits):
        """dlogits [H,B,K] contiguous -> (dx [B,F], grads in self.params order)"""
        lib = _lib.lib()
        H = self.H
        x = st["x"]
        B, F = x.shape
        dev = x.device
        K = self.lin2[0].out_features
        has_bias = self.lin2[0].bias is not None
        db2 = None
        if has_bias:
            db2 = torch.empty((H, K), dtype=torch.float32, device=dev)
            _lib.check(lib.selavi_heads_colsum(_lib.ptr(dlogits), _lib.ptr(db2), H, B, K, _stream()), "selavi_heads_colsum")
        dx = torch.empty((B, F), dtype=torch.float32, device=dev)
        if self.mlp:
            Fh = self.lin1[0].out_features
            a1, z1 = st["a1"], st["z1"]
            # dW2[h] = dlogits[h]^T @ a1[h]      [K, Fh]
            dW2 = torch.empty((H, K, Fh), dtype=torch.float32, device=dev)
            _bgemm(H, K, Fh, B, A=dlogits, a=(B * K, 1, K), B=a1, b=(B * Fh, Fh, 1), C=dW2, c=(K * Fh, Fh, 1))
            # da1[h] = dlogits[h] @ W2[h]        [B, Fh]
            da1 = torch.empty((H, B, Fh), dtype=torch.float32, device=dev)
            _bgemm(H, B, Fh, K, A=dlogits, a=(B * K, K, 1), B_tbl=st["w2"], b=(0, Fh, 1), C=da1, c=(B * Fh, Fh, 1))
            sums = torch.empty((H, 2, Fh), dtype=torch.float64, device=dev)
            _lib.check(lib.selavi_heads_bn_bwd_reduce(_lib.ptr(da1), _lib.ptr(st["m2"]), _lib.ptr(z1), _lib.ptr(st["scale"]),
                                                      _lib.ptr(st["shift"]), _lib.ptr(st["mean"]), _lib.ptr(st["invstd"]), H,
                                                      B, Fh, _lib.ptr(sums), _stream()), "selavi_heads_bn_bwd_reduce")
            dgamma = sums[:, 1].float()
            dbeta = sums[:, 0].float()
            bn0 = self.bn[0]
            if _world(bn0) > 1:
                sums = sums.clone()
                _allreduce(sums, bn0)
            dz1 = torch.empty_like(z1)
            _lib.check(lib.selavi_heads_bn_bwd_apply(_lib.ptr(da1), _lib.ptr(st["m2"]), _lib.ptr(z1), _lib.ptr(st["scale"]),
                                                     _lib.ptr(st["shift"]), _lib.ptr(st["mean"]), _lib.ptr(st["invstd"]),
                                                     _lib.ptr(sums), st["count"], H, B, Fh, _lib.ptr(dz1), _stream()),
                       "selavi_heads_bn_bwd_apply")
            # dW1[h] = dz1[h]^T @ (x * m1[h])    [Fh, F]
            dW1 = torch.empty((H, Fh, F), dtype=torch.float32, device=dev)
            if st["m1"] is None:
                _bgemm(H, Fh, F, B, A=dz1, a=(B * Fh, 1, Fh), B=x, b=(0, F, 1), C=dW1, c=(Fh * F, F, 1))
            else:
                self._dw1_masked(dz1, x, st["m1"], dW1, H, B, Fh, F)
            # dd1[h] = dz1[h] @ W1[h]  [B, F];  dx = sum_h dd1[h] * m1[h]
            dd1 = torch.empty((H, B, F), dtype=torch.float32, device=dev)
            _bgemm(H, B, F, Fh, A=dz1, a=(B * Fh, Fh, 1), B_tbl=st["w1"], b=(0, F, 1), C=dd1, c=(B * F, F, 1))
            _lib.check(lib.selavi_heads_sum_masked(_lib.ptr(dd1), _lib.ptr(st["m1"]), _lib.ptr(dx), H, B * F, 0, _stream()),
                       "selavi_heads_sum_masked")
            grads = []
            for h in range(H):
                grads += [dW1[h], dgamma[h].contiguous(), dbeta[h].contiguous(), dW2[h]] + ([db2[h]] if has_bias else [])
        else:
            dW2 = torch.empty((H, K, F), dtype=torch.float32, device=dev)
            if st["m1"] is None:
                _bgemm(H, K, F, B, A=dlogits, a=(B * K, 1, K), B=x, b=(0, F, 1), C=dW2, c=(K * F, F, 1))
            else:
                self._dw1_masked(dlogits, x, st["m1"], dW2, H, B, K, F)
            dd = torch.empty((H, B, F), dtype=torch.float32, device=dev)
            _bgemm(H, B, F, K, A=dlogits, a=(B * K, K, 1), B_tbl=st["w2"], b=(0, F, 1), C=dd, c=(B * F, F, 1))
            _lib.check(lib.selavi_heads_sum_masked(_lib.ptr(dd), _lib.ptr(st["m1"]), _lib.ptr(dx), H, B * F, 0, _stream()),
                       "selavi_heads_sum_masked")
            grads = []
            for h in range(H):
                grads += [dW2[h]] + ([db2[h]] if has_bias else [])
        return dx, grads

    @staticmethod
    def _dw1_masked(dz, x, m1, dW, H, B, Fh, F):
        """dW[h] = dz[h]^T @ (x * m1[h]): the masked operand must be the reduced-over (k = batch row) matrix B(k, n),
        so swap roles: dW[h]^T = (x*m1[h])^T @ dz[h]  ->  write with transposed C strides."""
        _bgemm(H, F, Fh, B, A=x, a=(0, 1, F), Amask=m1, am_bs=B * F, B=dz, b=(B * Fh, Fh, 1), C=dW, c=(Fh * F, 1, F))


class _HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run, x, *params):
        logits, st = run.forward(x, save=True)
        ctx.run, ctx.st = run, st
        outs = tuple(logits[h] for h in range(run.H))
        ctx.shape = tuple(logits.shape)
        return outs

    @staticmethod
    def backward(ctx, *douts):
        run = ctx.run
        H, B, K = ctx.shape
        dl = torch.empty((H, B, K), dtype=torch.float32, device=ctx.st["x"].device)
        for h, d in enumerate(douts):
            if d is None:
                dl[h].zero_()
            else:
                dl[h].copy_(d)
        dx, grads = run.backward(ctx.st, dl)
        ctx.st = None
        return (None, dx) + tuple(grads)


def heads_forward(heads, x):
    """All heads of one modality in batched launches (replaces the per-head loop model.py:233-252)."""
    if not (x.is_cuda and x.dtype == torch.float32):
        raise ValueError("selavi_b200 heads need float32 CUDA input (no CPU fallback)")
    if x.dim() != 2:
        x = x.reshape(x.shape[0], -1)
    run = _HeadsRun(heads, heads[0].training)
    need = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in run.params))
    if need:
        return list(_HeadsFn.apply(run, x, *run.params))
    with torch.no_grad():
        logits, _ = run.forward(x, save=False)
    return [logits[h] for h in range(run.H)]


# ====================================================================================================== loss
class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, targets, H, *logits):
        lib = _lib.lib()
        B, K = logits[0].shape
        dev = logits[0].device
        logits = [l.contiguous() for l in logits]
        # the heads of a modality are slices of ONE [H,B,K] buffer (heads_forward): address them by stride, which needs
        # no device pointer table (building one is a synchronising host-to-device copy in the middle of the step)
        base, stride, tbl = logits[0], B * K, None
        if any(l.data_ptr() != base.data_ptr() + h * stride * 4 for h, l in enumerate(logits)):
            tbl = _ptr_table(logits)
        targets = targets.to(torch.int64)
        if H == 1 and targets.dim() == 1:
            sb, sh = targets.stride(0), 0
        else:
            sb, sh = targets.stride(0), targets.stride(1)
        rows = torch.empty((H, B), dtype=torch.float32, device=dev)
        mean = torch.empty(1, dtype=torch.float32, device=dev)
        dl = torch.empty((H, B, K), dtype=torch.float32, device=dev)
        _lib.check(lib.selavi_ce_loss(_lib.ptr(tbl), _lib.ptr(base), stride, _lib.ptr(targets), sb, sh, H, B, K, 1.0 / (H * B), _lib.ptr(rows),
                                      _lib.ptr(mean), _lib.ptr(dl), _stream()), "selavi_ce_loss")
        ctx.dl = dl
        ctx.keep = (logits, tbl, targets)
        return mean.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        dl = ctx.dl * dloss
        return (None, None) + tuple(dl[h] for h in range(dl.shape[0]))


def get_loss(activations, targets, headcount=1):
    """utils.py:377-387: mean over heads of cross_entropy(activations[h], targets[:, h])."""
    if headcount == 1:
        acts = [activations] if torch.is_tensor(activations) else list(activations)
    else:
        acts = list(activations)[:headcount]
    if not acts[0].is_cuda:
        raise ValueError("selavi_b200.get_loss needs CUDA logits (no CPU fallback)")
    return _CEFn.apply(targets, len(acts), *acts)
