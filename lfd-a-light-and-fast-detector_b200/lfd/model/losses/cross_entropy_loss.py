# -*- coding: utf-8 -*-
"""CrossEntropyLoss -- API of lfd/model/losses/cross_entropy_loss.py:12-50 (inside LFD.get_loss the softmax
cross entropy and its gradient are computed by lfd_detection_loss)."""
import torch.nn as nn
import torch.nn.functional as F

from .utils import weight_reduce_loss

__all__ = ['CrossEntropyLoss']


class CrossEntropyLoss(nn.Module):

    def __init__(self, reduction='mean', loss_weight=1.0):
        super(CrossEntropyLoss, self).__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        loss = F.cross_entropy(cls_score, label, reduction='none')
        if weight is not None:
            weight = weight.float()
        return self.loss_weight * weight_reduce_loss(loss, weight=weight, reduction=reduction, avg_factor=avg_factor)
