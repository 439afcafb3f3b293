# -*- coding: utf-8 -*-
"""SmoothL1Loss / MSELoss / BCEWithLogitsLoss / QualityFocalLoss -- constructor and call signatures of
lfd/model/losses/{smooth_l1_loss.py:47-96, mse_loss.py:16-50, bce_with_logits_loss.py:47-71, gfocal_loss.py:79-141}.
Inside LFD.get_loss these losses and their gradients are evaluated by lfd_detection_loss (csrc/losses.cu); the stand-alone modules
keep the reference's element-wise tensor formulas (API shims, not on the hot path)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils import weight_reduce_loss

__all__ = ['SmoothL1Loss', 'MSELoss', 'BCEWithLogitsLoss', 'QualityFocalLoss']


class _Pointwise(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super(_Pointwise, self).__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def _reduce(self, loss, weight, avg_factor, reduction_override):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction_override if reduction_override else self.reduction, avg_factor)


class SmoothL1Loss(_Pointwise):
    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super(SmoothL1Loss, self).__init__(reduction, loss_weight)
        assert beta > 0
        self.beta = beta

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        d = (pred - target).abs()
        return self._reduce(torch.where(d < self.beta, 0.5 * d * d / self.beta, d - 0.5 * self.beta), weight, avg_factor, reduction_override)


class MSELoss(_Pointwise):
    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        return self._reduce((pred - target) ** 2, weight, avg_factor, reduction_override)


class BCEWithLogitsLoss(_Pointwise):
    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        if cls_score.dim() != label.dim():          # class indices -> one-hot rows (index num_classes = background = all zero)
            onehot = label.new_zeros((label.shape[0], cls_score.shape[-1] + 1), dtype=cls_score.dtype)
            onehot.scatter_(1, label.clamp(min=0).view(-1, 1), 1.0)
            label = onehot[:, :cls_score.shape[-1]]
            if weight is not None:
                weight = weight.view(-1, 1).expand(-1, cls_score.shape[-1])
        if weight is not None:
            weight = weight.float()
        return self._reduce(F.binary_cross_entropy_with_logits(cls_score, label.float(), reduction='none'), weight, avg_factor, reduction_override)


class QualityFocalLoss(_Pointwise):
    def __init__(self, use_sigmoid=True, beta=2.0, reduction='mean', loss_weight=1.0):
        super(QualityFocalLoss, self).__init__(reduction, loss_weight)
        assert use_sigmoid is True, 'Only sigmoid in QFL supported now.'
        self.use_sigmoid, self.beta = use_sigmoid, beta

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        label, score = target
        sig = pred.sigmoid()
        loss = F.binary_cross_entropy_with_logits(pred, torch.zeros_like(pred), reduction='none') * sig.pow(self.beta)
        pos = torch.nonzero((label >= 0) & (label < pred.shape[1])).squeeze(1)
        pl = label[pos].long()
        loss[pos, pl] = F.binary_cross_entropy_with_logits(pred[pos, pl], score[pos], reduction='none') * (score[pos] - sig[pos, pl]).abs().pow(self.beta)
        return self._reduce(loss.sum(dim=1), weight, avg_factor, reduction_override)
