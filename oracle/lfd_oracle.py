# -*- coding: utf-8 -*-
"""CPU oracle for the LFD dense-conv hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker / CPU baseline.

It restates, in plain torch fp32 / numpy on the CPU, the algorithm of the
reference (YonghaoHe/LFD-A-Light-and-Fast-Detector, paths relative to the
reference root):

    forward            lfd/model/lfd.py:511-542
      backbone         lfd/model/backbone/lfd_resnet.py:354-439 (stem), :96-154 (FasterBlock),
                       :21-93 (FastBlock), :157-215 (FastestBlock), :441-501 (stages, taps)
      neck             lfd/model/neck/simple_neck.py:35-74
      head             lfd/model/head/lfd_head.py:85-185
    point coordinates  lfd/model/lfd.py:84-107
    label assignment   lfd/model/lfd.py:109-259
    get_loss           lfd/model/lfd.py:284-395
    focal loss         lfd/model/losses/build/sigmoid_focal_loss/src/cuda/sigmoid_focal_loss_cuda.cu:24-97
    IoU loss           lfd/model/losses/iou_loss.py:11-123
    cross entropy      lfd/model/losses/cross_entropy_loss.py:12-22
    reductions         lfd/model/losses/utils.py:28-54
    decode / results   lfd/model/lfd.py:261-282, :397-509, :544-655
    NMS                lfd/model/utils/nms.py:7-59,119-220 ; build/nms/src/cpu/nms_cpu.cpp:8-66

Pinning: tests/golden/*.pt were produced by tests/gen_golden.py, which imports
the reference's own modules from /root/reference (CPU, stubs for pycuda /
data_pipeline / native exts) and runs them on the same seeded weights and
inputs; tests/test_oracle_vs_golden.py checks every function here against
those vectors and against the reference's docstring known-answer vectors.
The sigmoid focal loss has no CPU implementation in the reference; its
restatement is pinned against torchvision.ops.sigmoid_focal_loss instead.

Forward flavours:
  * forward(..., emulate=None): the reference arithmetic (fp32).
  * forward(..., emulate='bf16' | 'fp16'): same graph with roundings to the CUDA
    path's 16-bit storage type inserted at exactly the points where the CUDA
    path stores 16-bit values (see DESIGN.md "rounding points") -- the Gate-B
    oracle of SURVEY.md section 7.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default (lfd_resnet.py:10-18 builds it with defaults)
GN_EPS = 1e-5  # nn.GroupNorm default


# ----------------------------------------------------------------------------------------------
# model configurations (the `prepare_model()` blocks of the shipped config scripts)
# ----------------------------------------------------------------------------------------------
def _cfg(stem_mode, stem_channels, arch, channels, out_indices, num_classes, ranges, cls_loss, merge, assign, block_mode='faster', head_k=1,
         head_norm=True):
    return dict(
        backbone=dict(block_mode=block_mode, stem_mode=stem_mode, input_channels=3, stem_channels=stem_channels,
                      body_architecture=list(arch), body_channels=list(channels), out_indices=tuple(out_indices)),
        neck=dict(num_neck_channels=128),
        head=dict(num_classes=num_classes, num_head_channels=128, num_conv_layers=2, gn_groups=16, conv_kernel_size=head_k, norm=head_norm,
                  share_head_flag=True, merge_path_flag=merge, classification_loss_type=cls_loss,
                  regression_loss_type='IoULoss'),
        lfd=dict(num_classes=num_classes, regression_ranges=tuple(ranges), gray_range_factors=(0.9, 1.1),
                 range_assign_mode=assign, distance_to_bbox_mode='sigmoid'),
    )


_WF_RANGES = ((4, 20), (20, 40), (40, 80), (80, 160), (160, 320))
_TT_RANGES = ((4, 32), (32, 64), (64, 128), (128, 256))
_TL_RANGES = ((4, 32), (32, 64), (64, 128), (128, 256), (256, 512))
CONFIGS = {
    # WIDERFACE_train/WIDERFACE_LFD_{XS,S,M,L}.py:76-158
    'WIDERFACE_XS': _cfg('faster', 32, [4, 2, 2, 3], [64, 64, 64, 64], ((0, 3), (1, 1), (2, 1), (3, 0), (3, 2)), 1, _WF_RANGES, 'FocalLoss', True, 'dist'),
    'WIDERFACE_S': _cfg('faster', 64, [4, 2, 2, 3], [64, 64, 64, 128], ((0, 3), (1, 1), (2, 1), (3, 0), (3, 2)), 1, _WF_RANGES, 'FocalLoss', True, 'dist'),
    'WIDERFACE_M': _cfg('fast', 64, [3, 2, 1, 1, 1], [64, 64, 64, 128, 128], ((0, 2), (1, 1), (2, 0), (3, 0), (4, 0)), 1, _WF_RANGES, 'FocalLoss', True, 'dist'),
    'WIDERFACE_L': _cfg('fast', 64, [4, 2, 2, 1, 1], [64, 64, 64, 128, 128], ((0, 3), (1, 1), (2, 1), (3, 0), (4, 0)), 1, _WF_RANGES, 'FocalLoss', True, 'dist'),
    # TT100K_train/TT100K_LFD_{L,S}.py
    'TT100K_L': _cfg('fast', 64, [5, 3, 2, 2], [64, 64, 128, 128], ((0, 4), (1, 2), (2, 1), (3, 1)), 45, _TT_RANGES, 'CrossEntropyLoss', False, 'longer'),
    'TT100K_S': _cfg('faster', 64, [4, 2, 1, 1], [64, 64, 64, 128], ((0, 3), (1, 1), (2, 0), (3, 0)), 45, _TT_RANGES, 'CrossEntropyLoss', False, 'longer'),
    # TrafficLight_train/TL_LFD_L.py:95-150: the head has NO norm layers (conv + bias + ReLU towers)
    'TL_L': _cfg('fast', 64, [5, 3, 2, 2, 2], [64, 64, 128, 128, 128], ((0, 4), (1, 2), (2, 1), (3, 1), (4, 1)), 1, _TL_RANGES, 'FocalLoss', True, 'dist',
                 head_norm=False),
    # not shipped: small nets that exercise the remaining block / stem modes (lfd_resnet.py:21-93 FastBlock, :157-215 FastestBlock,
    # 'fastest' stem :420-439) and conv_kernel_size=3 head towers (lfd_head.py:47-49)
    'TEST_FAST': _cfg('fast', 64, [2, 1, 1], [64, 64, 128], ((0, 1), (1, 0), (2, 0)), 2, ((4, 32), (32, 64), (64, 128)), 'FocalLoss', True, 'dist',
                      block_mode='fast', head_k=3),
    'TEST_FASTEST': _cfg('fastest', 32, [2, 2], [64, 64], ((0, 1), (1, 1)), 1, ((8, 64), (64, 128)), 'FocalLoss', False, 'dist', block_mode='fastest'),
}


def strides_of(cfg):
    """lfd_resnet.py:297-304 -- stride is per *stage*, so two taps of one stage share a stride."""
    bb = cfg['backbone']
    stem_stride = 2 if bb['stem_mode'] == 'fast' else 4
    return [stem_stride * 2 ** (s + 1) for (s, _) in sorted(bb['out_indices'])]


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def fp16r(t):
    return t.to(torch.float16).to(torch.float32)


_ROUNDERS = {None: None, False: None, True: bf16r, 'bf16': bf16r, 'fp16': fp16r}


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
class _Net(object):
    """Walks the reference module graph from a state_dict (reference key names)."""

    def __init__(self, cfg, sd, emulate, trace=None):
        self.cfg, self.sd = cfg, {k: v.detach().float() for k, v in sd.items()}
        self.rnd = _ROUNDERS[emulate]     # rounding to the CUDA path's 16-bit storage type (None: reference fp32 arithmetic)
        self.emu = self.rnd is not None
        self.trace = trace  # optional dict: conv key -> stored NCHW output of that fused layer (debugging aid)

    def r(self, t):
        return self.rnd(t) if self.emu else t

    def w(self, key):
        w = self.sd[key]
        return self.rnd(w) if self.emu else w

    def conv(self, x, key, stride, pad):
        b = self.sd.get(key + '.bias')
        return F.conv2d(x, self.w(key + '.weight'), None, stride=stride, padding=pad), b

    def bn(self, y, key, bias):
        sd = self.sd
        scale = sd[key + '.weight'] / torch.sqrt(sd[key + '.running_var'] + BN_EPS)
        shift = sd[key + '.bias'] - sd[key + '.running_mean'] * scale
        if bias is not None:
            shift = shift + bias * scale
        return y * scale[None, :, None, None] + shift[None, :, None, None]

    def conv_bn(self, x, ckey, nkey, stride, pad, relu, residual=None):
        if self.emu:   # CUDA path: scale folded into the weights before the bf16 rounding, shift added as bf16 (rounding point Rw)
            sd = self.sd
            scale = sd[nkey + '.weight'] / torch.sqrt(sd[nkey + '.running_var'] + BN_EPS)
            shift = sd[nkey + '.bias'] - sd[nkey + '.running_mean'] * scale
            if sd.get(ckey + '.bias') is not None:
                shift = shift + sd[ckey + '.bias'] * scale
            y = F.conv2d(x, self.rnd(sd[ckey + '.weight'] * scale[:, None, None, None]), None, stride=stride, padding=pad)
            y = y + self.rnd(shift)[None, :, None, None]
        else:
            y, b = self.conv(x, ckey, stride, pad)
            y = self.bn(y, nkey, b)
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        y = self.r(y)
        if self.trace is not None:
            self.trace[ckey] = y
        return y

    def stem(self, x):
        mode = self.cfg['backbone']['stem_mode']
        p = '_backbone._stem.'
        if mode == 'fast':  # lfd_resnet.py:356-376
            spec = [(0, 1, 3, 2), (3, 4, 1, 1)]
        elif mode == 'faster':  # :378-418
            spec = [(0, 1, 3, 2), (3, 4, 1, 1), (6, 7, 3, 2), (9, 10, 1, 1)]
        else:  # 'fastest' :420-439
            spec = [(0, 1, 3, 2), (3, 4, 3, 2)]
        for (ci, ni, k, s) in spec:
            x = self.conv_bn(x, p + '%d' % ci, p + '%d' % ni, s, k // 2, True)
        return x

    def block(self, x, prefix, first):
        mode = self.cfg['backbone']['block_mode']
        stride = 2 if first else 1
        identity = x
        if first:  # lfd_resnet.py:458-469 : conv1x1 s2 + BN, no activation
            identity = self.conv_bn(x, prefix + '._downsample.0', prefix + '._downsample.1', 2, 0, False)
        if mode == 'faster':  # :135-154
            out = self.conv_bn(x, prefix + '._conv1', prefix + '._norm1', stride, 1, True)
            out = self.conv_bn(out, prefix + '._conv2', prefix + '._norm2', 1, 1, True, residual=identity)
        elif mode == 'fast':  # :70-93
            out = self.conv_bn(x, prefix + '._conv1', prefix + '._norm1', stride, 1, True)
            out = self.conv_bn(out, prefix + '._conv2', prefix + '._norm2', 1, 0, True)
            out = self.conv_bn(out, prefix + '._conv3', prefix + '._norm3', 1, 1, True, residual=identity)
        else:  # fastest :196-215
            out = self.conv_bn(x, prefix + '._conv1', prefix + '._norm1', stride, 1, True)
            out = self.conv_bn(out, prefix + '._conv2', prefix + '._norm2', 1, 1, True, residual=identity)
        return out

    def backbone(self, x):
        bb = self.cfg['backbone']
        x = self.stem(x)
        outs = []
        taps = sorted(bb['out_indices'])
        max_stage = max(t[0] for t in taps)
        for i, nb in enumerate(bb['body_architecture'][:max_stage + 1]):
            for j in range(nb):
                x = self.block(x, '_backbone.stage%d.%d' % (i, j), j == 0)
                if (i, j) in taps:
                    outs.append(x)
        return outs

    def gn_relu(self, y, key, groups):
        """nn.GroupNorm + ReLU.  In emulation mode the statistics are those of the STORED (bf16) conv output, i.e.
        GroupNorm is applied to the tensor that exists in memory (rounding point Rg)."""
        n, c, h, w = y.shape
        ys = self.r(y).reshape(n, groups, -1)
        yg = ys.double()
        mean = yg.mean(dim=2)
        var = yg.var(dim=2, unbiased=False)
        rstd = (1.0 / torch.sqrt(var + GN_EPS)).float()
        mean = mean.float()
        out = ((ys - mean[:, :, None]) * rstd[:, :, None]).reshape(n, c, h, w)
        out = out * self.sd[key + '.weight'][None, :, None, None] + self.sd[key + '.bias'][None, :, None, None]
        return self.r(F.relu(out))

    def tower(self, x, prefix):
        hd = self.cfg['head']
        k = hd.get('conv_kernel_size', 1)
        step = 3 if hd.get('norm', True) else 2          # conv, [norm], activation (lfd_head.py:85-106)
        for i in range(hd['num_conv_layers']):
            y, b = self.conv(x, prefix + '.%d' % (step * i), 1, k // 2)
            if hd.get('norm', True):
                assert b is None  # bias=False when a norm follows (lfd_head.py:98)
                x = self.gn_relu(y, prefix + '.%d' % (step * i + 1), hd['gn_groups'])
            else:                 # no norm: conv + bias + ReLU; the CUDA path adds the bias as a bf16 row on the tensor core
                x = self.r(F.relu(y + (self.r(b) if self.emu else b)[None, :, None, None]))
            if self.trace is not None:
                self.trace[prefix + '.%d' % (step * i) + ':raw'] = self.r(y)
                self.trace[prefix + '.%d' % (step * i) + ':act'] = x
        return x

    def head_level(self, x, l):
        hd = self.cfg['head']
        nl = hd['num_conv_layers']
        p = '_head.head%d_' % l
        if hd['merge_path_flag']:  # lfd_head.py:91-106, final convs at index 0 of cls/reg paths
            t = self.tower(x, p + 'merge_path')
            cls_in, reg_in, last = t, t, 0
        else:  # :108-135, towers live inside the cls/reg paths, final conv at index 3*nl
            cls_in = self.tower(x, p + 'classification_path')
            reg_in = self.tower(x, p + 'regression_path')
            last = (3 if hd.get('norm', True) else 2) * nl
        ck, rk = p + 'classification_path.%d' % last, p + 'regression_path.%d' % last
        cls = F.conv2d(cls_in, self.w(ck + '.weight'), self.sd[ck + '.bias'])
        reg = F.conv2d(reg_in, self.w(rk + '.weight'), self.sd[rk + '.bias'])
        reg = reg * self.sd['_head._scales.%d._scale' % l]  # lfd_head.py:177-180 (scale multiplies the bias too)
        return cls, reg

    def forward(self, x):
        x = self.r(x.float())
        feats = self.backbone(x)
        cls_list, reg_list, sizes = [], [], []
        for l, f in enumerate(feats):
            # neck: simple_neck.py:35-47 (1x1 conv + BN + ReLU)
            nk = self.conv_bn(f, '_neck.neck%d.0' % l, '_neck.neck%d.1' % l, 1, 0, True)
            cls, reg = self.head_level(nk, l)
            n, c, h, w = cls.shape
            sizes.append((h, w))
            cls_list.append(cls.permute(0, 2, 3, 1).reshape(n, h * w, c))  # lfd.py:526-540
            reg_list.append(reg.permute(0, 2, 3, 1).reshape(n, h * w, 4))
        return torch.cat(cls_list, 1), torch.cat(reg_list, 1), sizes


def forward(cfg, state_dict, x, emulate_bf16=False, trace=None, emulate=None):
    """-> (cls [N,P,C'], reg [N,P,4], [(H_l, W_l)])  C' = C (sigmoid/focal) or C+1 (cross entropy).
    emulate: None (reference fp32), 'bf16' or 'fp16' (roundings of the CUDA path's storage type inserted at its rounding
    points); emulate_bf16=True is the older spelling of emulate='bf16'."""
    if emulate is None and emulate_bf16:
        emulate = 'bf16'
    with torch.no_grad():
        return _Net(cfg, state_dict, emulate, trace).forward(x)


# ----------------------------------------------------------------------------------------------
# point coordinates, label assignment
# ----------------------------------------------------------------------------------------------
def point_coordinates(sizes, strides):
    """lfd.py:84-107: points sit at the cell ORIGIN (x*s, y*s); row-major (y outer, x inner)."""
    out = []
    for (h, w), s in zip(sizes, strides):
        ys, xs = np.meshgrid(np.arange(h, dtype=np.int64) * s, np.arange(w, dtype=np.int64) * s, indexing='ij')
        out.append(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1))
    return out


def gray_ranges_of(ranges, factors=(0.9, 1.1)):
    lo, hi = min(factors), max(factors)
    return [(int(a * lo), int(b * hi)) for (a, b) in ranges]  # lfd.py:49-50


def assign_targets(cfg, sizes, gt_bboxes, gt_labels):
    """Scalar-per-(point, gt) restatement of lfd.py:155-259 for ONE image (numpy fp32).

    gt_bboxes [G,4] xywh float32, gt_labels [G] int64 ->
      cls_target [P,C] float32 in {-1, 0, (0,1]}, reg_target [P,4] float32, pos mask [P] bool.
    Deterministic tie rule (the reference's is order-unspecified, SURVEY 8a'): highest score wins,
    ties go to the lowest gt index.  reg_target rows of non-positive points are zero (the reference
    leaves unspecified values there; they are never read).
    """
    lc = cfg['lfd']
    strides = strides_of(cfg)
    ranges = lc['regression_ranges']
    grays = gray_ranges_of(ranges, lc['gray_range_factors'])
    mode = lc['range_assign_mode']
    C = lc['num_classes']
    pts = point_coordinates(sizes, strides)
    P = sum(p.shape[0] for p in pts)
    px = np.concatenate([p[:, 0] for p in pts]).astype(np.float32)
    py = np.concatenate([p[:, 1] for p in pts]).astype(np.float32)
    half = np.concatenate([np.full(p.shape[0], s / 2.0, np.float32) for p, s in zip(pts, strides)])
    lo = np.concatenate([np.full(p.shape[0], r[0], np.float32) for p, r in zip(pts, ranges)])
    hi = np.concatenate([np.full(p.shape[0], r[1], np.float32) for p, r in zip(pts, ranges)])
    glo = np.concatenate([np.full(p.shape[0], r[0], np.float32) for p, r in zip(pts, grays)])
    ghi = np.concatenate([np.full(p.shape[0], r[1], np.float32) for p, r in zip(pts, grays)])

    cls_t = np.zeros((P, C), np.float32)
    gray_any = np.zeros((P, C), bool)
    best = np.zeros(P, np.float32)  # best green score so far (0 = none)
    reg_t = np.zeros((P, 4), np.float32)
    one = np.float32(1.0)
    gt_bboxes = np.asarray(gt_bboxes, np.float32).reshape(-1, 4)
    for g in range(gt_bboxes.shape[0]):
        x, y, w, h = [np.float32(v) for v in gt_bboxes[g]]
        lab = int(gt_labels[g])
        cx = x + w / np.float32(2.0)
        cy = y + h / np.float32(2.0)
        xs = np.abs(px - cx) / half
        xs = np.where(xs >= 1, xs, one)
        xs = np.sqrt(one / xs)
        ysc = np.abs(py - cy) / half
        ysc = np.where(ysc >= 1, ysc, one)
        ysc = np.sqrt(one / ysc)
        score = (xs * ysc).astype(np.float32)
        d0 = px - x
        d1 = py - y
        d2 = ((x + w) - one) - px
        d3 = ((y + h) - one) - py
        delta = np.stack([d0, d1, d2, d3], -1).astype(np.float32)
        if mode == 'longer':
            measure = np.full(P, max(w, h), np.float32)
        elif mode == 'shorter':
            measure = np.full(P, min(w, h), np.float32)
        else:  # 'dist' -- per (point, gt) pair
            measure = delta.max(-1)
        hit = delta.min(-1) >= 0
        green = (lo <= measure) & (measure <= hi) & hit
        gray = (((glo <= measure) & (measure < lo)) | ((hi < measure) & (measure <= ghi))) & hit
        cls_t[:, lab] = np.where(green, np.maximum(cls_t[:, lab], score), cls_t[:, lab])
        gray_any[:, lab] |= gray
        better = green & (score > best)
        reg_t[better] = delta[better]
        best = np.where(better, score, best)
    cls_t[gray_any] = -1.0  # gray overrides green for that class (lfd.py:248-251)
    return cls_t, reg_t


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def sigmoid_focal_loss_forward(logits, targets, gamma, alpha):
    """sigmoid_focal_loss_cuda.cu:24-59.  logits [M,C] fp32, targets [M] int64 (== C means background)."""
    x = logits.float()
    M, C = x.shape
    d = torch.arange(C)[None, :]
    t = targets[:, None]
    c1 = (t == d).float()
    c2 = ((t >= 0) & (t != d)).float()
    p = 1.0 / (1.0 + torch.exp(-x))
    flt_min = torch.finfo(torch.float32).tiny
    term1 = torch.pow(1.0 - p, gamma) * torch.log(torch.clamp(p, min=flt_min))
    pos = (x >= 0).float()
    term2 = torch.pow(p, gamma) * (-1.0 * x * pos - torch.log(1.0 + torch.exp(x - 2.0 * x * pos)))
    return -c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)


def sigmoid_focal_loss_backward(logits, targets, d_losses, gamma, alpha):
    """sigmoid_focal_loss_cuda.cu:62-97."""
    x = logits.float()
    M, C = x.shape
    d = torch.arange(C)[None, :]
    t = targets[:, None]
    c1 = (t == d).float()
    c2 = ((t >= 0) & (t != d)).float()
    p = 1.0 / (1.0 + torch.exp(-x))
    flt_min = torch.finfo(torch.float32).tiny
    term1 = torch.pow(1.0 - p, gamma) * (1.0 - p - (p * gamma * torch.log(torch.clamp(p, min=flt_min))))
    pos = (x >= 0).float()
    log1mp = -1.0 * x * pos - torch.log(1.0 + torch.exp(x - 2.0 * x * pos))
    term2 = torch.pow(p, gamma) * (log1mp * (1.0 - p) * gamma - p)
    return (-c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)) * d_losses


def bbox_overlaps_aligned(b1, b2, eps=1e-6):
    """iou_loss.py:66-80,98-102 (is_aligned=True, mode='iou'): no +1, union clamped to eps."""
    lt = torch.max(b1[:, :2], b2[:, :2])
    rb = torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[:, 0] * wh[:, 1]
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = torch.clamp(a1 + a2 - overlap, min=eps)
    return overlap / union


def bbox_overlaps(b1, b2, eps=1e-6):
    """iou_loss.py:82-102 (is_aligned=False)."""
    if b1.shape[0] * b2.shape[0] == 0:
        return b1.new_zeros((b1.shape[0], b2.shape[0]))
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = torch.clamp(a1[:, None] + a2 - overlap, min=eps)
    return overlap / union


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """losses/utils.py:28-54."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def distance2bbox(points, distance, max_shape=None):
    """lfd.py:261-282 (clamp is inclusive of W and H)."""
    x1 = points[:, 0] - distance[:, 0]
    y1 = points[:, 1] - distance[:, 1]
    x2 = points[:, 0] + distance[:, 2]
    y2 = points[:, 1] + distance[:, 3]
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1])
        y1 = y1.clamp(min=0, max=max_shape[0])
        x2 = x2.clamp(min=0, max=max_shape[1])
        y2 = y2.clamp(min=0, max=max_shape[0])
    return torch.stack([x1, y1, x2, y2], -1)


def get_loss(cfg, cls_pred, reg_pred, sizes, annotation_batch, gamma=2.0, alpha=0.25, iou_eps=1e-6):
    """lfd.py:284-395 for FocalLoss|CrossEntropyLoss + IoULoss, distance_to_bbox_mode in {sigmoid, exp}.

    cls_pred/reg_pred may require grad (autograd gives the oracle gradients; the focal term uses the
    analytic backward of the .cu through a custom Function).  Returns dict(loss, classification_loss,
    regression_loss, n_pos, n_valid).
    """
    lc, hd = cfg['lfd'], cfg['head']
    C = lc['num_classes']
    strides = strides_of(cfg)
    N = cls_pred.shape[0]
    cls_targets, reg_targets = [], []
    for (bboxes, labels) in annotation_batch:
        ct, rt = assign_targets(cfg, sizes, bboxes, labels)
        cls_targets.append(torch.from_numpy(ct))
        reg_targets.append(torch.from_numpy(rt))
    cls_t = torch.stack(cls_targets).reshape(-1, C)
    reg_t = torch.stack(reg_targets).reshape(-1, 4)
    Cp = C + 1 if hd['classification_loss_type'] == 'CrossEntropyLoss' else C
    fc = cls_pred.reshape(-1, Cp)
    fr = reg_pred.reshape(-1, 4)
    green = torch.where(cls_t.min(dim=-1)[0] >= 0)[0]
    fc, fr, cls_t, reg_t = fc[green], fr[green], cls_t[green], reg_t[green]
    max_scores, max_idx = cls_t.max(dim=-1)
    pos = torch.where(max_scores >= 0.001)[0]
    label = max_idx * (max_scores >= 0.001) + C * (max_scores < 0.001)
    n_pos = int(pos.numel())
    if hd['classification_loss_type'] == 'FocalLoss':
        el = _FocalFn.apply(fc, label, gamma, alpha)
    else:
        el = F.cross_entropy(fc, label, reduction='none')
    cls_loss = weight_reduce_loss(el, None, 'mean', n_pos + 1)
    frp, rtp = fr[pos], reg_t[pos]
    if n_pos > 0:
        pts = torch.from_numpy(np.concatenate(point_coordinates(sizes, strides), 0)).repeat(N, 1)[green][pos].float()
        tgt = distance2bbox(pts, rtp)
        if lc['distance_to_bbox_mode'] == 'exp':
            pred = distance2bbox(pts, frp.float().exp())
        else:
            his = torch.cat([torch.full((h * w,), float(max(r)), dtype=torch.float32)
                             for (h, w), r in zip(sizes, lc['regression_ranges'])]).repeat(N)[green][pos]
            pred = distance2bbox(pts, frp.sigmoid() * his[:, None])
        ious = bbox_overlaps_aligned(pred, tgt, 1e-6).clamp(min=iou_eps)
        reg_loss = weight_reduce_loss(-ious.log(), None, 'mean', n_pos)
    else:
        reg_loss = frp.sum()
    return dict(loss=cls_loss + reg_loss, classification_loss=cls_loss, regression_loss=reg_loss,
                n_pos=n_pos, n_valid=int(green.numel()))


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t, gamma, alpha):
        ctx.save_for_backward(x, t)
        ctx.g, ctx.a = gamma, alpha
        return sigmoid_focal_loss_forward(x, t, gamma, alpha)

    @staticmethod
    def backward(ctx, d):
        x, t = ctx.saved_tensors
        return sigmoid_focal_loss_backward(x, t, d, ctx.g, ctx.a), None, None, None


# ----------------------------------------------------------------------------------------------
# post-process
# ----------------------------------------------------------------------------------------------
def nms(dets, thr):
    """Greedy NMS of nms_cpu.cpp:8-66 on dets [n,5] float32 (x1,y1,x2,y2,score).

    Returns int64 indices into dets, in kept (score-descending) order.  Suppression is strict
    `iou > thr`, IoU has no +1 and no epsilon.  Sort is made stable (index ascending on equal
    scores); the reference's is unspecified on ties.
    """
    dets = np.asarray(dets, np.float32).reshape(-1, 5)
    n = dets.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    x1, y1, x2, y2, sc = [dets[:, i] for i in range(5)]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-sc, kind='stable')
    suppressed = np.zeros(n, bool)
    keep = []
    thr = np.float32(thr)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def multiclass_nms(boxes, scores, score_thr, iou_thr, class_agnostic=False, max_num=-1, nms_fn=None):
    """nms.py:161-220 + batched_nms :119-158.  boxes [K,4], scores [K,C] (bg column already dropped).

    Returns (dets [k,5], labels [k], src [k]) with src = flat index point*C + class of each kept row.
    The class-offset arithmetic (label * (max_coordinate + 1) added in fp32) is reproduced exactly.
    """
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    scores = np.asarray(scores, np.float32)
    K, C = scores.shape
    b = np.repeat(boxes[:, None, :], C, axis=1).reshape(-1, 4)
    s = scores.reshape(-1)
    lab = np.tile(np.arange(C, dtype=np.int64), K)
    inds = np.nonzero(s > np.float32(score_thr))[0]
    b, s, lab = b[inds], s[inds], lab[inds]
    if inds.size == 0:
        return np.zeros((0, 5), np.float32), np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    if class_agnostic:
        bn = b
    else:
        off = lab.astype(np.float32) * (b.max() + np.float32(1))
        bn = (b + off[:, None]).astype(np.float32)
    keep = (nms_fn or nms)(np.concatenate([bn, s[:, None]], 1), iou_thr)   # nms_fn: e.g. the reference's compiled nms_ext.nms
    out = np.concatenate([bn[keep], s[keep, None]], 1)
    if not class_agnostic:
        out[:, :4] = out[:, :4] - off[keep][:, None]
    if max_num > 0:
        out, keep = out[:max_num], keep[:max_num]
    return out.astype(np.float32), lab[keep], inds[keep]


def decode_image(cfg, cls_i, reg_i, sizes, height, width, resize_scale=1.0):
    """lfd.py:434-499 up to the NMS call: scores [P,C] and boxes [P,4] (fp32, torch CPU)."""
    lc, hd = cfg['lfd'], cfg['head']
    strides = strides_of(cfg)
    pts = point_coordinates(sizes, strides)
    split = [p.shape[0] for p in pts]
    cs, rs = cls_i.float().split(split, 0), reg_i.float().split(split, 0)
    sc_all, bx_all = [], []
    for l in range(len(split)):
        if hd['classification_loss_type'] == 'CrossEntropyLoss':
            sc = cs[l].softmax(dim=1)[:, :-1]
        else:
            sc = cs[l].sigmoid()
        p = torch.from_numpy(pts[l])
        if lc['distance_to_bbox_mode'] == 'exp':
            d = rs[l].exp()
        else:
            d = rs[l].sigmoid() * float(max(lc['regression_ranges'][l]))
        bx_all.append(distance2bbox(p, d, max_shape=(height, width)))
        sc_all.append(sc)
    boxes = torch.cat(bx_all) / resize_scale
    return torch.cat(sc_all), boxes


def get_results(cfg, cls, reg, sizes, meta_batch, score_thr, iou_thr, class_agnostic=False, nms_fn=None):
    """lfd.py:397-432: per image rows [label, score, x, y, w, h] with w = x2 - x1 + 1.

    Also returns, per image, the flat source indices (point*C + class) of the kept rows.
    """
    results, srcs = [], []
    for i in range(cls.shape[0]):
        m = meta_batch[i]
        sc, bx = decode_image(cfg, cls[i], reg[i], sizes, m['resized_height'], m['resized_width'], m['resize_scale'])
        dets, labels, src = multiclass_nms(bx.numpy(), sc.numpy(), score_thr, iou_thr, class_agnostic, nms_fn=nms_fn)
        rows = []
        for d, lab in zip(dets, labels):
            rows.append([int(lab), float(d[4]), float(d[0]), float(d[1]),
                         float(np.float32(d[2] - d[0]) + np.float32(1)), float(np.float32(d[3] - d[1]) + np.float32(1))])
        results.append(rows)
        srcs.append(src)
    return results, srcs


def normalize_image_u8(image_hwc_u8):
    """augmentation_pipeline.py:31-36 `simple_normalize`: (x/255 - 0.5)/0.5 on BGR uint8 -> float32 HWC.

    Parity unpinned by reference execution (albumentations is not installed); formula restated from
    albumentations.Normalize(mean=.5, std=.5, max_pixel_value=255): (img - mean*255) * (1/(std*255)).
    """
    img = np.asarray(image_hwc_u8).astype(np.float32)
    return ((img - np.float32(0.5 * 255.0)) * np.float32(1.0 / (0.5 * 255.0))).astype(np.float32)
