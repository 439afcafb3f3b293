# -*- coding: utf-8 -*-
"""IoULoss -- API of lfd/model/losses/iou_loss.py:11-123,286-321.  Inside LFD.get_loss the decode + IoU loss +
gradient are one fused kernel (lfd_detection_loss); this module keeps the stand-alone call signature."""
import torch
import torch.nn as nn

from .utils import weighted_loss

__all__ = ['IoULoss', 'bbox_overlaps']


def bbox_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False, eps=1e-6):
    assert mode in ['iou', 'iof']
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if is_aligned:
        assert rows == cols
    if rows * cols == 0:
        return bboxes1.new(rows, 1) if is_aligned else bboxes1.new(rows, cols)
    a, b = (bboxes1, bboxes2) if is_aligned else (bboxes1[:, None, :], bboxes2[None, :, :])
    wh = (torch.min(a[..., 2:], b[..., 2:]) - torch.max(a[..., :2], b[..., :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    area1 = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    union = area1 + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - overlap if mode == 'iou' else area1 + 0 * overlap
    return overlap / torch.max(union, union.new_tensor([eps]))


@weighted_loss
def iou_loss(pred, target, eps=1e-6):
    return -bbox_overlaps(pred, target, is_aligned=True).clamp(min=eps).log()


class IoULoss(nn.Module):

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super(IoULoss, self).__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != 'none'):
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        return self.loss_weight * iou_loss(pred, target, weight, eps=self.eps, reduction=reduction, avg_factor=avg_factor, **kwargs)
