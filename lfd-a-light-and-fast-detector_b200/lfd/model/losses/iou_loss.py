# -*- coding: utf-8 -*-
"""IoULoss / GIoULoss / DIoULoss / CIoULoss -- API of lfd/model/losses/iou_loss.py:105-430.  Inside LFD.get_loss the decode + box
loss + gradient are one fused kernel (lfd_detection_loss); the stand-alone modules here evaluate the element-wise loss and its
gradient w.r.t. the predicted boxes in liblfd_b200.so (lfd_box_loss).  CUDA only."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _native as nat
from .utils import weight_reduce_loss

__all__ = ['IoULoss', 'GIoULoss', 'DIoULoss', 'CIoULoss', 'bbox_overlaps']


def bbox_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False, eps=1e-6):
    """lfd/model/losses/iou_loss.py:11-102 (plain tensor arithmetic: an API helper, not on the hot path)."""
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


class _BoxLossFunction(Function):
    """element-wise loss [n] of box pairs [n,4] with the gradient w.r.t. the predictions precomputed by the same kernel."""

    @staticmethod
    def forward(ctx, pred, target, kind, eps):
        if not pred.is_cuda:
            raise RuntimeError('lfd_b200 box losses are CUDA only (no CPU fallback)')
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        n = p.shape[0]
        loss = torch.empty((n,), dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        with torch.cuda.device(p.device):
            nat.check(nat.lib().lfd_box_loss(int(kind), nat.ptr(p), nat.ptr(t), n, float(eps), nat.ptr(loss), nat.ptr(grad), nat.stream_ptr()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        (grad,) = ctx.saved_tensors
        return grad * d_loss[:, None], None, None, None


class _BoxLoss(nn.Module):
    KIND = nat.REG_IOU

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super(_BoxLoss, self).__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != 'none'):
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        loss = _BoxLossFunction.apply(pred, target, self.KIND, self.eps)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


class IoULoss(_BoxLoss):       # -log(max(IoU, eps)), reference :105-123,286-321
    KIND = nat.REG_IOU


class GIoULoss(_BoxLoss):      # reference :125-170,324-357
    KIND = nat.REG_GIOU


class DIoULoss(_BoxLoss):      # reference :173-224,360-394
    KIND = nat.REG_DIOU


class CIoULoss(_BoxLoss):      # reference :227-283,397-430
    KIND = nat.REG_CIOU
