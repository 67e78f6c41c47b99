"""Generator training step of BASELINE.json configs[4] (SURVEY.md 8d cfg 5):

    G forward (train-mode BatchNorm) -> DAMSM word/sentence ranking loss on the final image (frozen image and text
    encoders) + MSE against the real image pyramid -> backward -> gradient all-reduce (NCCL) -> Adam(2e-4, (0.5, 0.999)).

The reference ships no SR training loop or optimiser; this composes its loss functions
(miscc/losses.py:499-522, :779-784) with its published hyper-parameters.  Everything on the generator side runs on
libtgsr_sm100 kernels (tgsr_b200/train.py, losses.py, csrc/optim.cu).  The frozen Inception-v3 image encoder
(CNN_ENCODER, util.py:262-368) is out of the hot path's scope and runs on torchvision/cuDNN as in the reference.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import losses
from ._lib import bump_weights_epoch, call, ptr


class FlatBucket:
    """All trainable parameters of `modules` re-homed into ONE flat fp32 buffer, with a same-shaped flat gradient
    buffer that autograd accumulates into directly (p.grad are views).  One all-reduce and one Adam launch per step."""

    def __init__(self, modules: Sequence[nn.Module]):
        self.params: List[nn.Parameter] = []
        seen = set()
        for m in modules:
            for p in m.parameters():
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p)); self.params.append(p)
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n); n += (p.numel() + 3) // 4 * 4          # 16-byte aligned slices
        self.numel = n
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + p.numel()].view(p.shape)
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):                       # re-attach if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def all_reduce(self, group=None) -> None:
        """Sum the flat gradient over the ranks (NCCL over NVLink on GPUs; gloo in the CPU tests).  The 1/world
        average is folded into the optimiser's grad_scale."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)


class FusedAdam:
    """torch.optim.Adam semantics on a FlatBucket, one kernel per step (csrc/optim.cu).

    The step count lives on the device (`step_t`, one int32) as the single source of truth, so eager steps, captured
    steps and CUDA-graph replays can be mixed: the bias corrections are computed in-kernel from it.  `t` mirrors it on
    the host for reporting only (replays advance it through `note_replay`)."""

    def __init__(self, bucket: FlatBucket, lr: float = 2e-4, betas=(0.5, 0.999), eps: float = 1e-8):
        self.b, self.lr, self.betas, self.eps, self.t = bucket, lr, betas, eps, 0
        self.m = torch.zeros_like(bucket.flat); self.v = torch.zeros_like(bucket.flat)
        self.step_t = torch.zeros(1, device=bucket.flat.device, dtype=torch.int32)

    def step(self, grad_scale: float = 1.0) -> None:
        """One update; safe both eagerly and inside a CUDA-graph capture (no host-side constant changes between steps)."""
        b = self.b
        self.step_t += 1
        if not torch.cuda.is_current_stream_capturing():
            self.t += 1
        call('tgsr_adam_devstep_f32', b.flat.device, ptr(b.flat), ptr(b.grad), ptr(self.m), ptr(self.v), C.c_size_t(b.numel), self.lr,
             self.betas[0], self.betas[1], self.eps, ptr(self.step_t), float(grad_scale))
        bump_weights_epoch()                   # parameters were written through raw pointers: packed / folded caches are stale

    step_capturable = step

    def note_replay(self) -> None:
        """A captured step was replayed: the device counter advanced by itself; keep the host mirror and the caches in step."""
        self.t += 1
        bump_weights_epoch()


import os as _os
_FUSED_CONV_RELU = int(_os.environ.get('TGSR_ENCODER_FUSED_CONV', '1'))
_OWN_MAXPOOL = int(_os.environ.get('TGSR_ENCODER_OWN_MAXPOOL', '1'))
_OWN_RELU_BWD = int(_os.environ.get('TGSR_ENCODER_OWN_RELU_BWD', '1'))


def _relu_bwd(gy, y):
    """gy * (y > 0).  gy is often a channel slice of a concatenation's gradient (pixels strided): csrc/pool16.cu reads it with
    16-byte vectors where the library falls back to a scalar strided kernel."""
    B, C, H, W = y.shape
    st = gy.stride()
    if (_OWN_RELU_BWD and gy.dtype == torch.bfloat16 and C % 8 == 0 and st[1] == 1 and st[3] % 8 == 0 and st[2] == W * st[3]
            and st[0] == H * st[2] and gy.data_ptr() % 16 == 0 and y.is_contiguous(memory_format=torch.channels_last)
            and y.stride()[1] == 1):
        from ._lib import call, ptr, FMT_BF16
        import ctypes as C_
        out = torch.empty_like(y)
        call('tgsr_relu_bwd_nhwc16', y.device, ptr(gy), C_.c_longlong(st[3]), ptr(y), ptr(out), C_.c_longlong(B * H * W), C, FMT_BF16)
        return out
    return torch.ops.aten.threshold_backward(gy, y, 0)


class _ConvBiasReluFn(torch.autograd.Function):
    """relu(conv2d(x, w) + b) of the frozen image encoder as cuDNN's fused conv-bias-activation (one kernel instead of conv,
    broadcast add and clamp); only the input gradient exists (the encoder's weights are frozen)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        y = torch.cudnn_convolution_relu(x, w, b, stride, padding, (1, 1), 1)
        ctx.save_for_backward(x, w, y)
        ctx.conf = (tuple(stride), tuple(padding))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, padding = ctx.conf
        g = _relu_bwd(gy, y)
        dx = torch.ops.aten.convolution_backward(g, x, w, None, stride, padding, (1, 1), False, (0, 0), 1, (True, False, False))[0]
        return dx, None, None, None, None


class _MaxPool3s2Fn(torch.autograd.Function):
    """F.max_pool2d(x, 3, stride 2) on channels-last bf16 tensors through csrc/pool16.cu (gather kernels both ways)."""

    @staticmethod
    def forward(ctx, x):
        from ._lib import call, ptr, FMT_BF16
        x = x.contiguous(memory_format=torch.channels_last)
        B, C, H, W = x.shape
        Ho, Wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        idx = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.uint8)
        call('tgsr_maxpool3s2_nhwc16', x.device, ptr(x), ptr(y), ptr(idx), B, H, W, C, FMT_BF16)
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        from ._lib import call, ptr, FMT_BF16
        idx, = ctx.saved_tensors
        B, C, H, W = ctx.shape
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty((B, C, H, W), device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
        call('tgsr_maxpool3s2_nhwc16_bwd', gy.device, ptr(gy), ptr(idx), ptr(gx), B, H, W, C, FMT_BF16)
        return gx


def _max_pool_3s2(x):
    if x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and _OWN_MAXPOOL:
        return _MaxPool3s2Fn.apply(x)
    return F.max_pool2d(x, kernel_size=3, stride=2)


class FrozenImageEncoder(nn.Module):
    """CNN_ENCODER (util.py:262-368) with random-init Inception-v3 weights (no network for the pretrained file), frozen,
    eval mode.  Library code (torchvision + cuDNN) outside the hot path: only its input gradient is used."""

    def __init__(self, nef: int = 256, seed: int = 100):
        super().__init__()
        from torchvision.models import inception_v3
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        m = inception_v3(weights=None, aux_logits=True, init_weights=True)
        for n in ('Conv2d_1a_3x3', 'Conv2d_2a_3x3', 'Conv2d_2b_3x3', 'Conv2d_3b_1x1', 'Conv2d_4a_3x3', 'Mixed_5b', 'Mixed_5c',
                  'Mixed_5d', 'Mixed_6a', 'Mixed_6b', 'Mixed_6c', 'Mixed_6d', 'Mixed_6e', 'Mixed_7a', 'Mixed_7b', 'Mixed_7c'):
            setattr(self, n, getattr(m, n))
        self.emb_features = nn.Conv2d(768, nef, 1, bias=False)
        self.emb_cnn_code = nn.Linear(2048, nef)
        self.emb_features.weight.data.uniform_(-0.1, 0.1); self.emb_cnn_code.weight.data.uniform_(-0.1, 0.1)
        torch.random.set_rng_state(g)
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def forward(self, x):
        if self.autocast_bf16 and x.is_cuda:
            # library code outside the hot path: let cuDNN run it the way a mixed-precision inference stack would -- eval-mode
            # BatchNorm folded into the conv weights (the encoder is frozen), bf16 channels-last tensors end to end -- so that
            # it does not dominate the bf16 step; features return as fp32
            self._fold_bf16(x.device)
            f, c = self._forward(x, torch.bfloat16)
            return f.float().contiguous(), c.float()
        return self._forward(x)

    autocast_bf16 = False
    _folded_on = None

    def _load_from_state_dict(self, *a, **k):
        self._folded_on = None
        return super()._load_from_state_dict(*a, **k)

    def _fold_bf16(self, device):
        """conv(bias=False) -> BatchNorm(eval) -> ReLU of every torchvision BasicConv2d becomes one bf16 conv with bias + ReLU:
        w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps).  The folded tensors live outside the
        state dict (rebuilt after load_state_dict or a device move)."""
        from ._lib import tensor_key
        key = (device, tensor_key(self.Conv2d_1a_3x3.conv.weight, False), tensor_key(self.emb_features.weight, False))
        if self._folded_on == key:                           # rebuilt after load_state_dict / invalidate_caches() / a device move
            return
        import types
        from torchvision.models.inception import BasicConv2d

        from torchvision.models.inception import InceptionA, InceptionB, InceptionC, InceptionD, InceptionE

        def folded(m, x):
            if x.dtype != torch.bfloat16:                    # the fp32 path keeps conv -> BatchNorm -> ReLU
                return BasicConv2d.forward(m, x)
            if x.is_cuda and _FUSED_CONV_RELU:
                return _ConvBiasReluFn.apply(x, m._fw, m._fb, m.conv.stride, m._fpad)
            return F.relu(F.conv2d(x, m._fw, m._fb, m.conv.stride, m._fpad), inplace=True)

        # branch wiring of the blocks whose pooling branch is avg_pool2d(3, stride 1, pad 1) -> 1x1 conv: that pair is ONE 3x3
        # conv with every tap = w / 9 (count_include_pad), which runs on tensor cores instead of two bandwidth-bound pooling
        # kernels (forward + backward: 3.5 ms of the step before)
        chains = {InceptionA: (('branch1x1',), ('branch5x5_1', 'branch5x5_2'), ('branch3x3dbl_1', 'branch3x3dbl_2', 'branch3x3dbl_3')),
                  InceptionC: (('branch1x1',), ('branch7x7_1', 'branch7x7_2', 'branch7x7_3'),
                               ('branch7x7dbl_1', 'branch7x7dbl_2', 'branch7x7dbl_3', 'branch7x7dbl_4', 'branch7x7dbl_5'))}

        def block_ac(blk, x):
            if x.dtype != torch.bfloat16:
                return type(blk).forward(blk, x)
            outs = []
            for chain in chains[type(blk)]:
                y = x
                for name in chain:
                    y = getattr(blk, name)(y)
                outs.append(y)
            outs.append(blk.branch_pool(x))
            return torch.cat(outs, 1)

        def block_b(blk, x):                                    # grid reduction: conv s2 | double 3x3 s2 | max pool
            if x.dtype != torch.bfloat16:
                return InceptionB.forward(blk, x)
            return torch.cat((blk.branch3x3(x), blk.branch3x3dbl_3(blk.branch3x3dbl_2(blk.branch3x3dbl_1(x))), _max_pool_3s2(x)), 1)

        def block_d(blk, x):
            if x.dtype != torch.bfloat16:
                return InceptionD.forward(blk, x)
            y = x
            for name in ('branch7x7x3_1', 'branch7x7x3_2', 'branch7x7x3_3', 'branch7x7x3_4'):
                y = getattr(blk, name)(y)
            return torch.cat((blk.branch3x3_2(blk.branch3x3_1(x)), y, _max_pool_3s2(x)), 1)

        def block_e(blk, x):
            if x.dtype != torch.bfloat16:
                return InceptionE.forward(blk, x)
            a = blk.branch3x3_1(x)
            b = blk.branch3x3dbl_2(blk.branch3x3dbl_1(x))
            return torch.cat((blk.branch1x1(x), blk.branch3x3_2a(a), blk.branch3x3_2b(a), blk.branch3x3dbl_3a(b),
                              blk.branch3x3dbl_3b(b), blk.branch_pool(x)), 1)

        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, BasicConv2d):
                    s = m.bn.weight / torch.sqrt(m.bn.running_var + m.bn.eps)
                    m._fw = (m.conv.weight * s.view(-1, 1, 1, 1)).to(device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
                    m._fb = (m.bn.bias - m.bn.running_mean * s).to(device, torch.bfloat16)
                    m._fpad = m.conv.padding
                    m.forward = types.MethodType(folded, m)
            for blk in self.modules():
                if isinstance(blk, (InceptionA, InceptionC, InceptionE)):
                    bp = blk.branch_pool
                    s = bp.bn.weight / torch.sqrt(bp.bn.running_var + bp.bn.eps)
                    w3 = (bp.conv.weight * s.view(-1, 1, 1, 1) / 9.0).expand(-1, -1, 3, 3)
                    bp._fw = w3.to(device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
                    bp._fpad = (1, 1)
                    blk.forward = types.MethodType(block_e if isinstance(blk, InceptionE) else block_ac, blk)
                elif isinstance(blk, InceptionB):
                    blk.forward = types.MethodType(block_b, blk)
                elif isinstance(blk, InceptionD):
                    blk.forward = types.MethodType(block_d, blk)
            self._ef_w = self.emb_features.weight.to(device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
            self._ec_w = self.emb_cnn_code.weight.to(device, torch.bfloat16)
            self._ec_b = self.emb_cnn_code.bias.to(device, torch.bfloat16)
        self._folded_on = key

    def _forward(self, x, dtype=None):
        x = F.interpolate(x, size=(299, 299), mode='bilinear', align_corners=False)
        if dtype is not None:
            x = x.to(dtype).contiguous(memory_format=torch.channels_last)
        x = self.Conv2d_2b_3x3(self.Conv2d_2a_3x3(self.Conv2d_1a_3x3(x)))
        x = _max_pool_3s2(x)
        x = self.Conv2d_4a_3x3(self.Conv2d_3b_1x1(x))
        x = _max_pool_3s2(x)
        x = self.Mixed_5d(self.Mixed_5c(self.Mixed_5b(x)))
        x = self.Mixed_6e(self.Mixed_6d(self.Mixed_6c(self.Mixed_6b(self.Mixed_6a(x)))))
        features = x
        x = self.Mixed_7c(self.Mixed_7b(self.Mixed_7a(x)))
        x = F.avg_pool2d(x, kernel_size=8).view(x.size(0), -1)
        if dtype is not None:
            return F.conv2d(features, self._ef_w), F.linear(x, self._ec_w, self._ec_b)
        return self.emb_features(features), self.emb_cnn_code(x)


class GeneratorTrainStep:
    def __init__(self, netGL: nn.Module, netGH: nn.Module, text_encoder: nn.Module, image_encoder: nn.Module, lr: float = 2e-4,
                 betas=(0.5, 0.999), group=None, conv_precision: str = 'tf32'):
        """conv_precision: 'fp32' (exact CUDA-core convs), 'tf32' (mma.sync TF32, what PyTorch's cuDNN convs compute by
        default) or 'bf16' (BASELINE.json configs[4]: tcgen05 convs on blocked bf16 activations with fp32 master weights,
        fp32 BatchNorm statistics and fp32 weight-gradient accumulation; tgsr_b200/train16.py)."""
        from . import set_precision, train
        self.conv_precision = conv_precision
        if conv_precision == 'bf16':
            set_precision('bf16')
            if hasattr(image_encoder, 'autocast_bf16'):
                image_encoder.autocast_bf16 = True
                image_encoder.to(memory_format=torch.channels_last)
            train.set_conv_precision('tf32')          # the fp32 edges (3-channel convs) and the weight gradients
        else:
            set_precision('fp32')
            train.set_conv_precision(conv_precision)  # TF32 tensor-core convs by default, as PyTorch's cuDNN convs
        self.netGL, self.netGH, self.text_encoder, self.image_encoder = netGL, netGH, text_encoder, image_encoder
        netGL.train(); netGH.train(); text_encoder.eval(); image_encoder.eval()
        for p in netGL.ca_net.parameters():                                 # c_code is unused by the SR nets (model.py:51-52)
            p.requires_grad = False
        self.bucket = FlatBucket([netGL, netGH])
        self.opt = FusedAdam(self.bucket, lr, betas)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1:
            self.sync_replicas()
        self.ev = None
        self._lab = None
        self._graph = None

    def sync_replicas(self) -> None:
        """Rank 0's parameters and BatchNorm buffers become every rank's (what DDP does at construction).  BatchNorm
        running statistics are per-replica during training, as in plain DDP; call this again before evaluating or
        checkpointing from a rank other than 0."""
        dist.broadcast(self.bucket.flat, 0, group=self.group)
        for net in (self.netGL, self.netGH):
            for buf in net.buffers():
                if buf.is_floating_point() or buf.dtype == torch.int64:
                    dist.broadcast(buf, 0, group=self.group)
        bump_weights_epoch()

    def _fwd_bwd(self, LR, real, captions, cap_lens, tmax: int):
        """Zero the flat gradient, forward (train-mode BN), losses, backward: everything before the exchange step."""
        B = LR.shape[0]
        from . import get_precision, set_precision
        want = 'bf16' if self.conv_precision == 'bf16' else 'fp32'
        if get_precision() != want:
            set_precision(want)
        self.bucket.zero_grad()
        with torch.no_grad():
            words, sent = self.text_encoder._forward_tmax(captions, cap_lens, tmax)
            mask = (captions == 0)[:, :tmax]
        fake_imgL, _att, _mu, _logvar = self.netGL(LR, sent, words, mask)
        fine, _a, _one = self.netGH(LR, fake_imgL, LR)
        labels = self._labels(B, LR.device)
        errG, _ = losses.generator_lossor_wordsentOnly(self.image_encoder, fine, words, sent, labels, cap_lens, None)
        mse = losses.MSE(list(fine) + list(fake_imgL), list(real) + list(real))
        total = errG + mse
        total.backward()
        return {'total': total.detach(), 'damsm': errG.detach(), 'mse': mse.detach()}

    def _exchange(self, time_allreduce: bool = False) -> None:
        if time_allreduce:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
        self.bucket.all_reduce(self.group)
        if time_allreduce:
            e.record(); self.ev = (s, e)

    def _update(self, capturable: bool) -> None:
        self.opt.step(1.0 / self.world)

    def _step_impl(self, LR, real, captions, cap_lens, tmax: int, capturable: bool, time_allreduce: bool = False):
        out = self._fwd_bwd(LR, real, captions, cap_lens, tmax)
        self._exchange(time_allreduce)
        self._update(capturable)
        return out

    def _labels(self, B, dev):
        if self._lab is None or self._lab.shape[0] != B:
            self._lab = torch.arange(B, device=dev)
        return self._lab

    def step(self, LR: torch.Tensor, real: Sequence[torch.Tensor], captions: torch.Tensor, cap_lens: torch.Tensor, tmax: int,
             time_allreduce: bool = False) -> Dict[str, torch.Tensor]:
        """LR [B,3,32,32]; real = [64x64, 128x128, 256x256] targets; captions int64 [B,18] sorted by length (desc)."""
        return self._step_impl(LR, real, captions, cap_lens, tmax, False, time_allreduce)

    # ------------------------------------------------------------------ steps fed with decoded images
    IMSIZE = (32, 64, 128, 256)        # cfg.TREE.BASE_SIZE * 2**i, BRANCH_NUM 4 (datasets.py:1575-1578)

    def _inputs_from_images(self, hr_u8: torch.Tensor, imsize=None):
        """uint8 [B, S, S, 3] decoded + cropped HR batch on the device -> (LR, real pyramid) with the reference's PIL
        arithmetic (datasets.py:112-148 `get_imgs`: ret[0] is the LR input, ret[1:] the MSE targets), csrc/pyramid.cu."""
        from . import pyramid
        ret = pyramid.real_pyramid(hr_u8, list(imsize or self.IMSIZE))
        return ret[0], ret[1:]

    def step_from_images(self, hr_u8: torch.Tensor, captions: torch.Tensor, cap_lens: torch.Tensor, tmax: int,
                         time_allreduce: bool = False) -> Dict[str, torch.Tensor]:
        """One step from decoded images: the input pyramid the reference builds per sample on the host (PIL, DataLoader
        workers) runs on the device in front of the step."""
        LR, real = self._inputs_from_images(hr_u8)
        return self._step_impl(LR, real, captions, cap_lens, tmax, False, time_allreduce)

    def capture_from_images(self, hr_u8, captions, cap_lens, tmax: int, warmup: int = 3) -> None:
        """As capture(), with the input pyramid inside the graph: the static inputs are the uint8 batch and the captions."""
        self._static_img = [hr_u8.clone(), captions.clone(), cap_lens.clone()]
        st = self._static_img

        def inputs():
            LR, real = self._inputs_from_images(st[0])
            return LR, real, st[1], st[2]
        self._graph_img, self._graph_img_b, self._graph_img_out, self.graph_img_kernels = self._capture_graphs(inputs, tmax, warmup)

    def step_graph_from_images(self, hr_u8, captions, cap_lens, time_allreduce: bool = False) -> Dict[str, torch.Tensor]:
        st = self._static_img
        if hr_u8.data_ptr() != st[0].data_ptr():
            st[0].copy_(hr_u8, non_blocking=True)
            st[1].copy_(captions, non_blocking=True); st[2].copy_(cap_lens, non_blocking=True)
        self._replay(self._graph_img, self._graph_img_b, time_allreduce)
        return self._graph_img_out

    # ------------------------------------------------------------------ whole-step CUDA graph
    def _capture_graphs(self, body_inputs, tmax: int, warmup: int):
        """Capture one step.  NOTE: the `warmup` eager steps that precede the capture (they size every workspace and warm the
        allocator) are REAL optimiser steps on the static batch: parameters, Adam moments and BatchNorm running statistics
        advance by `warmup` steps (the step counter counts them).
        On one GPU the whole step is ONE graph; with N > 1 ranks the NCCL all-reduce stays outside:
        graph A = zero-grad + forward + losses + backward, the eager all-reduce of the flat bucket, graph B = fused Adam
        (capturing the collective itself hung the ranks in round 1, so it is not attempted)."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd(*body_inputs(), tmax)
                self._exchange()
                self._update(True)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        lib = _lib.load()
        n0 = lib.tgsr_launch_count()
        g_a = torch.cuda.CUDAGraph()
        g_b = None
        if self.world == 1:
            with torch.cuda.graph(g_a):
                out = self._fwd_bwd(*body_inputs(), tmax)
                self._update(True)
        else:
            with torch.cuda.graph(g_a):
                out = self._fwd_bwd(*body_inputs(), tmax)
            g_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_b, pool=g_a.pool()):
                self._update(True)
        return g_a, g_b, out, int(lib.tgsr_launch_count() - n0)

    def _replay(self, g_a, g_b, time_allreduce: bool = False) -> None:
        g_a.replay()
        if g_b is not None:
            self._exchange(time_allreduce)
            g_b.replay()
        self.opt.note_replay()

    def capture(self, LR, real, captions, cap_lens, tmax: int, warmup: int = 3) -> None:
        """Capture forward + losses + backward (+ Adam) of one step into CUDA graphs (static shapes).  The ~1400 launches
        of a step are launch-bound from Python; replay removes that."""
        self._static = [LR.clone(), [r.clone() for r in real], captions.clone(), cap_lens.clone()]
        st = self._static
        self._graph, self._graph_b, self._graph_out, self.graph_kernels = self._capture_graphs(
            lambda: (st[0], st[1], st[2], st[3]), tmax, warmup)

    def step_graph(self, LR, real, captions, cap_lens, time_allreduce: bool = False) -> Dict[str, torch.Tensor]:
        st = self._static
        if LR.data_ptr() != st[0].data_ptr():
            st[0].copy_(LR, non_blocking=True)
            for d, r in zip(st[1], real):
                d.copy_(r, non_blocking=True)
            st[2].copy_(captions, non_blocking=True); st[3].copy_(cap_lens, non_blocking=True)
        self._replay(self._graph, self._graph_b, time_allreduce)
        return self._graph_out
