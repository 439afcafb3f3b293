# -*- coding: utf-8 -*-
"""nms / batched_nms / multiclass_nms with the call signatures of lfd/model/utils/nms.py:7-59,119-220, backed by
lfd_nms in liblfd_b200.so (bitonic sort + greedy sweep in one CTA; no host round trip).  CUDA tensors (or numpy
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


def batched_nms(bboxes, scores, inds, nms_cfg, class_agnostic=False):
    nms_cfg_ = nms_cfg.copy()
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        bboxes_for_nms = bboxes
    else:
        offsets = inds.to(bboxes) * (bboxes.max() + 1)
        bboxes_for_nms = bboxes + offsets[:, None]
    nms_type = nms_cfg_.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError('only nms_cfg type "nms" is on the LFD path (lfd.py:76)')
    nms_bboxes, kept = nms(torch.cat([bboxes_for_nms, scores[:, None]], -1), **nms_cfg_)
    if not class_agnostic:
        nms_bboxes[:, :4] = nms_bboxes[:, :4] - offsets[kept][:, None]
    return nms_bboxes, kept


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, 4)
    scores = multi_scores[:, :-1]
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    labels = torch.arange(num_classes, dtype=torch.long, device=scores.device).view(1, -1).expand_as(scores)
    bboxes, scores, labels = bboxes.reshape(-1, 4), scores.reshape(-1), labels.reshape(-1)
    inds = (scores > score_thr).nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if inds.numel() == 0:
        return bboxes, labels
    dets, keep = batched_nms(bboxes, scores, labels, nms_cfg)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, labels[keep]
