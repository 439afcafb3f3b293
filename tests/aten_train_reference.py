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


def _cnr(x, conv, norm, relu):
    x = conv(x)
    if norm is not None:
        x = norm(x)
    return F.relu(x) if relu else x


def train_forward(model, x):
    """x: float32 [N,3,H,W] on CUDA -> (classification [N,P,C'], regression [N,P,4]), differentiable; records the level
    sizes in model._head_indexes_to_feature_map_sizes like the eval path does."""
    bb, neck, head = model._backbone, model._neck, model._head
    for conv, norm, relu in bb.stem_layers():                       # reference backbone/lfd_resnet.py:354-439
        x = _cnr(x, conv, norm, relu)
    taps = list(bb._out_indices)
    feats = [None] * len(taps)
    for si, stage in enumerate(bb.stages()):                        # :441-473, blocks :96-154
        for bi, block in enumerate(stage):
            identity = x if block._downsample is None else block._downsample(x)
            pairs = block.conv_norm_pairs()
            for li, (conv, norm) in enumerate(pairs):
                x = conv(x)
                if norm is not None:
                    x = norm(x)
                if li == len(pairs) - 1:
                    x = x + identity
                x = F.relu(x)
            if (si, bi) in taps:
                feats[taps.index((si, bi))] = x
    cls_out, reg_out = [], []
    for l, f in enumerate(feats):
        conv, norm = neck.level(l)                                  # neck/simple_neck.py:35-47,67-74
        t = _cnr(f, conv, norm, True)
        cls_tower, reg_tower, fin_cls, fin_reg = head.level_paths(l)   # head/lfd_head.py:85-143,164-185
        tc = t
        for conv, norm in cls_tower:
            tc = _cnr(tc, conv, norm, True)
        if reg_tower is cls_tower:
            tr = tc
        else:
            tr = t
            for conv, norm in reg_tower:
                tr = _cnr(tr, conv, norm, True)
        c, r = fin_cls(tc), fin_reg(tr)
        if head.uses_scale:
            r = r * head._scales[l]._scale                            # Scale multiplies the bias too (lfd_head.py:177-180)
        n, _, h, w = c.shape
        model._head_indexes_to_feature_map_sizes[l] = (h, w)
        cls_out.append(c.permute(0, 2, 3, 1).reshape(n, h * w, -1))   # lfd.py:526-540
        reg_out.append(r.permute(0, 2, 3, 1).reshape(n, h * w, 4))
    return torch.cat(cls_out, 1), torch.cat(reg_out, 1)
