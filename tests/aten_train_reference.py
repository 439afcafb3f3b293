# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY (the checker of the native training plan, never imported by the product).

Evaluates the SAME module graph the native training plan is built from with ATen ops in fp32 NCHW -- i.e. the reference's own
arithmetic (lfd/model/lfd.py:511-542 and the modules it calls: backbone/lfd_resnet.py:96-154,354-473, neck/simple_neck.py:35-74,
head/lfd_head.py:85-185) -- and lets autograd differentiate it.  tests/test_gpu_train.py compares losses, output gradients and
parameter gradients of the hand-written CUDA path against this.
"""
import torch
import torch.nn.functional as F

__all__ = ['train_forward']


class _RoundBf16(torch.autograd.Function):
    """Rounding point of the native bf16 training path: the value is rounded to bf16 in the forward, and the gradient arriving at
    that tensor is rounded to bf16 in the backward (the native path stores both the tensor and its gradient as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _conv(x, conv, rnd):
    if rnd is None:
        return conv(x)
    w = conv.weight + (conv.weight.to(torch.bfloat16).float() - conv.weight).detach()     # bf16 operand, fp32 master gradient
    return rnd(F.conv2d(x, w, conv.bias, conv.stride, conv.padding))


def _cnr(x, conv, norm, relu, rnd=None):
    x = _conv(x, conv, rnd)
    if norm is not None:
        x = norm(x)
    x = F.relu(x) if relu else x
    return x if rnd is None else rnd(x)


def train_forward(model, x, emulate_bf16=False):
    """x: float32 [N,3,H,W] -> (classification [N,P,C'], regression [N,P,4]), differentiable; records the level
    sizes in model._head_indexes_to_feature_map_sizes like the eval path does.  emulate_bf16: the same graph with the native path's
    rounding points (input, conv weights, every stored conv output z, every stored layer output y, and the gradients of z / y)."""
    rnd = _RoundBf16.apply if emulate_bf16 else None
    bb, neck, head = model._backbone, model._neck, model._head
    if rnd is not None:
        x = x.to(torch.bfloat16).float()
    for conv, norm, relu in bb.stem_layers():                       # reference backbone/lfd_resnet.py:354-439
        x = _cnr(x, conv, norm, relu, rnd)
    taps = list(bb._out_indices)
    feats = [None] * len(taps)
    for si, stage in enumerate(bb.stages()):                        # :441-473, blocks :96-154
        for bi, block in enumerate(stage):
            if block._downsample is None:
                identity = x
            else:
                ds = list(block._downsample)
                identity = _cnr(x, ds[0], ds[1] if len(ds) > 1 else None, False, rnd)
            pairs = block.conv_norm_pairs()
            for li, (conv, norm) in enumerate(pairs):
                x = _conv(x, conv, rnd)
                if norm is not None:
                    x = norm(x)
                if li == len(pairs) - 1:
                    x = x + identity
                x = F.relu(x)
                if rnd is not None:
                    x = rnd(x)
            if (si, bi) in taps:
                feats[taps.index((si, bi))] = x
    cls_out, reg_out = [], []
    for l, f in enumerate(feats):
        conv, norm = neck.level(l)                                  # neck/simple_neck.py:35-47,67-74
        t = _cnr(f, conv, norm, True, rnd)
        cls_tower, reg_tower, fin_cls, fin_reg = head.level_paths(l)   # head/lfd_head.py:85-143,164-185
        tc = t
        for conv, norm in cls_tower:
            tc = _cnr(tc, conv, norm, True, rnd)
        if reg_tower is cls_tower:
            tr = tc
        else:
            tr = t
            for conv, norm in reg_tower:
                tr = _cnr(tr, conv, norm, True, rnd)

        def final(conv, t_):
            if rnd is None:
                return conv(t_)
            w = conv.weight + (conv.weight.to(torch.bfloat16).float() - conv.weight).detach()
            return F.conv2d(t_, w, conv.bias)                         # fp32 outputs
        c, r = final(fin_cls, tc), final(fin_reg, tr)
        if head.uses_scale:
            r = r * head._scales[l]._scale                            # Scale multiplies the bias too (lfd_head.py:177-180)
        n, _, h, w = c.shape
        model._head_indexes_to_feature_map_sizes[l] = (h, w)
        cls_out.append(c.permute(0, 2, 3, 1).reshape(n, h * w, -1))   # lfd.py:526-540
        reg_out.append(r.permute(0, 2, 3, 1).reshape(n, h * w, 4))
    return torch.cat(cls_out, 1), torch.cat(reg_out, 1)
