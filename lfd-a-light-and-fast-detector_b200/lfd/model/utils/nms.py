# -*- coding: utf-8 -*-
"""nms / batched_nms / multiclass_nms with the call signatures of lfd/model/utils/nms.py:7-59,119-220, backed by
lfd_nms / lfd_multiclass_nms in liblfd_b200.so (threshold, bitonic sort, class-offset suppression on the device).  CUDA tensors (or numpy
arrays with device_id) only -- there is no CPU NMS in lfd_b200.  soft_nms / nms_match are not on the LFD path."""
import ctypes as C

import numpy as np
import torch

from ... import _native as nat

__all__ = ['nms', 'batched_nms', 'multiclass_nms']


def _native_nms(dets, iou_thr):
    n = int(dets.shape[0])
    d = dets.detach().float().contiguous()
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=d.device)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=d.device)
    ws = torch.empty(nat.lib().lfd_nms_workspace_bytes(n), dtype=torch.uint8, device=d.device)
    with torch.cuda.device(d.device):
        nat.check(nat.lib().lfd_nms(nat.ptr(d), n, float(iou_thr), nat.ptr(ws), nat.ptr(keep), nat.ptr(n_keep), nat.stream_ptr()))
    return keep[:int(n_keep.item())]


def nms(dets, iou_thr, device_id=None):
    if isinstance(dets, torch.Tensor):
        is_numpy, dets_th = False, dets
    elif isinstance(dets, np.ndarray):
        is_numpy = True
        if device_id is None:
            raise RuntimeError('lfd_b200 has no CPU NMS: pass device_id or a CUDA tensor')
        dets_th = torch.from_numpy(dets).to('cuda:{}'.format(device_id))
    else:
        raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))
    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    else:
        if not dets_th.is_cuda:
            raise RuntimeError('lfd_b200 has no CPU NMS: move dets to a CUDA device')
        inds = _native_nms(dets_th, iou_thr)
    if is_numpy:
        inds = inds.cpu().numpy()
    return dets[inds, :], inds


def _device_nms(boxes, box_per_class, scores, score_stride, labels_in, n, num_classes, score_thr, iou_thr, class_agnostic):
    """lfd_multiclass_nms: candidates (threshold) + class-offset NMS in two launches.  -> (dets [k,5], labels [k] int64, rows [k] int64)."""
    dev = boxes.device
    total = n if labels_in is not None else n * num_classes
    cap = max(int(total), 1)
    L = nat.lib()
    ws = torch.empty(max(L.lfd_multiclass_nms_workspace_bytes(cap), 256), dtype=torch.uint8, device=dev)
    dets = torch.empty((cap, 5), dtype=torch.float32, device=dev)
    labels = torch.empty((cap,), dtype=torch.int32, device=dev)
    src = torch.empty((cap,), dtype=torch.int32, device=dev)
    count = torch.zeros((2,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(L.lfd_multiclass_nms(nat.ptr(boxes), int(box_per_class), nat.ptr(scores), int(score_stride), nat.ptr(labels_in), int(n), int(num_classes),
                                       float(score_thr), float(iou_thr), int(bool(class_agnostic)), cap, nat.ptr(ws), nat.ptr(dets), nat.ptr(labels),
                                       nat.ptr(src), nat.ptr(count), nat.ptr(count[1:]), nat.stream_ptr()))
    k = int(count[0].item())
    return dets[:k], labels[:k].long(), torch.div(src[:k].long(), num_classes, rounding_mode='floor')


def _nms_args(nms_cfg, class_agnostic=False):
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    if cfg.pop('type', 'nms') != 'nms':
        raise NotImplementedError('only nms_cfg type "nms" is on the LFD path (lfd.py:76)')
    return float(cfg.get('iou_thr', cfg.get('iou_threshold', 0.5))), class_agnostic


def batched_nms(bboxes, scores, inds, nms_cfg, class_agnostic=False):
    """reference :119-158: NMS that never suppresses across different `inds` (class labels).  -> (dets [k,5], keep [k]) with keep indexing the
    inputs, score-descending.  One native call: the label * (max coordinate + 1) offsets live inside the NMS kernel."""
    iou_thr, class_agnostic = _nms_args(nms_cfg, class_agnostic)
    if not bboxes.is_cuda:
        raise RuntimeError('lfd_b200 has no CPU NMS: move the boxes to a CUDA device')
    n = int(bboxes.shape[0])
    if n == 0:
        return torch.cat([bboxes, scores[:, None]], -1), inds.new_zeros(0, dtype=torch.long)
    labels_in = inds.to(torch.int32).contiguous()
    num_classes = int(labels_in.max().item()) + 1
    dets, _, rows = _device_nms(bboxes.detach().float().contiguous(), 0, scores.detach().float().contiguous(), 1, labels_in, n, num_classes,
                                0.0, iou_thr, class_agnostic)
    return dets.to(bboxes.dtype), rows


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """reference :161-220: per-class score threshold (strict >) + class-aware NMS.  multi_bboxes [n,4] or [n,C*4], multi_scores [n,C+1]
    (background last).  -> (dets [k,5], labels [k])."""
    iou_thr, class_agnostic = _nms_args(nms_cfg)
    if not multi_bboxes.is_cuda:
        raise RuntimeError('lfd_b200 has no CPU NMS: move the boxes to a CUDA device')
    n, num_classes = int(multi_scores.shape[0]), int(multi_scores.shape[1]) - 1
    scores = multi_scores.detach().float()
    if score_factors is not None:
        scores = torch.cat([scores[:, :-1] * score_factors[:, None], scores[:, -1:]], 1)
    scores = scores.contiguous()
    if n == 0:
        return multi_bboxes.new_zeros((0, 5)), multi_bboxes.new_zeros((0,), dtype=torch.long)
    dets, labels, _ = _device_nms(multi_bboxes.detach().float().contiguous(), int(multi_bboxes.shape[1] > 4), scores, num_classes + 1, None, n, num_classes,
                                  score_thr, iou_thr, class_agnostic)
    if max_num > 0:
        dets, labels = dets[:max_num], labels[:max_num]
    return dets.to(multi_bboxes.dtype), labels
