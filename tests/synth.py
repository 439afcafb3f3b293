# -*- coding: utf-8 -*-
"""Deterministic synthetic weights / inputs / annotations shared by tests/gen_golden.py (which feeds them to the
REFERENCE modules) and by the tests (which feed them to the oracle and to the CUDA path).

Every tensor is generated from a torch.Generator seeded by crc32(key), so values do not depend on module
construction order; a checksum of the generated state is stored in the golden files to catch RNG drift.
"""
import re
import zlib

import numpy as np
import torch


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def synth_state_dict(template, seed=666, cls_bias=-2.0):
    """template: mapping key -> tensor (shapes / dtypes), e.g. model.state_dict().  Shared head tensors
    (share_head_flag) are generated once under their head0 name."""
    out = {}
    for key, t in template.items():
        canon = re.sub(r'_head\.head\d+_', '_head.head0_', key)
        g = _gen(canon, seed)
        shape = tuple(t.shape)
        if key.endswith('num_batches_tracked'):
            v = torch.tensor(100, dtype=torch.long)
        elif key.endswith('running_mean'):
            v = torch.randn(shape, generator=g) * 0.2
        elif key.endswith('running_var'):
            v = torch.rand(shape, generator=g) + 0.5
        elif key.endswith('_scale'):
            v = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 4:  # conv weight: keep activations O(1) through the net
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * (1.6 / fan_in) ** 0.5
        elif key.endswith('.weight'):  # BN / GN gamma
            v = torch.rand(shape, generator=g) + 0.5
        elif key.endswith('.bias'):
            v = torch.randn(shape, generator=g) * 0.1
            if 'classification_path' in key and shape[0] <= 64:
                v = v + cls_bias  # so that only a small fraction of points passes the score threshold
        else:
            raise KeyError('synth_state_dict: unhandled key ' + key)
        out[key] = v.to(t.dtype)
    return out


def state_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


def synth_input(n, h, w, seed=666):
    g = _gen('input', seed)
    return torch.rand((n, 3, h, w), generator=g) * 2.0 - 1.0


def synth_image_u8(h, w, seed=666):
    g = _gen('image_u8', seed)
    return torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).numpy()


def synth_annotations(n, h, w, num_classes, seed=666, max_boxes=12):
    """[(bboxes float32 [G,4] xywh, labels int64 [G])]; image 1 (if any) is a negative image (G = 0).
    Box sides are log-uniform in [4, 320] to cover every regression range and the gray bands; coordinates
    use quarter-pixel fractions (not integers) so that exact ties between two boxes' scores are avoided."""
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        G = 0 if i == 1 else int(rng.randint(1, max_boxes + 1))
        boxes = np.zeros((G, 4), np.float32)
        for g in range(G):
            bw = float(np.exp(rng.uniform(np.log(4), np.log(min(320, w)))))
            bh = float(np.exp(rng.uniform(np.log(4), np.log(min(320, h)))))
            x = rng.uniform(0, max(w - bw, 1))
            y = rng.uniform(0, max(h - bh, 1))
            boxes[g] = [round(x * 4) / 4 + 0.13, round(y * 4) / 4 + 0.29, round(bw * 4) / 4 + 0.07, round(bh * 4) / 4 + 0.19]
        labels = rng.randint(0, num_classes, size=(G,)).astype(np.int64)
        out.append((boxes, labels))
    return out
