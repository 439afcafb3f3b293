# -*- coding: utf-8 -*-
"""LFDHead -- constructor / state_dict compatible with lfd/model/head/lfd_head.py:30-185
(LFDHeadV1 is not part of the hot path: no shipped config uses it)."""
import torch
import torch.nn as nn

from ..backbone.lfd_resnet import make_norm, make_activation

__all__ = ['LFDHead']


class Scale(nn.Module):
    def __init__(self, scale_factor=1.0):
        super(Scale, self).__init__()
        self._scale = nn.Parameter(torch.tensor(scale_factor, dtype=torch.float))


class LFDHead(nn.Module):

    def __init__(self, num_classes, num_input_channels, num_heads, num_head_channels=128, num_conv_layers=2,
                 conv_kernel_size=1, activation_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='BatchNorm2d'),
                 classification_loss_type='SmoothL1Loss', regression_loss_type='SmoothL1Loss', share_head_flag=False,
                 merge_path_flag=False):
        super(LFDHead, self).__init__()
        assert classification_loss_type in ['BCEWithLogitsLoss', 'FocalLoss', 'CrossEntropyLoss', 'QualityFocalLoss']
        assert regression_loss_type in ['SmoothL1Loss', 'MSELoss', 'IoULoss', 'GIoULoss', 'DIoULoss', 'CIoULoss']
        assert conv_kernel_size in [1, 3]
        self._num_classes = num_classes
        self._num_input_channels = num_input_channels
        self._num_head_channels = num_head_channels
        self._num_conv_layers = num_conv_layers
        self._conv_kernel_size = conv_kernel_size
        self._activation_cfg, self._norm_cfg = activation_cfg, norm_cfg
        self._share_head_flag, self._merge_path_flag = share_head_flag, merge_path_flag
        self._num_heads = num_heads
        self._classification_loss_type = classification_loss_type
        self._regression_loss_type = regression_loss_type
        if regression_loss_type in ['IoULoss', 'GIoULoss', 'DIoULoss', 'CIoULoss']:
            self._scales = nn.ModuleList([Scale(1.0) for _ in range(num_heads)])
        for i in range(num_heads):
            if i == 0 or not share_head_flag:
                paths = self._build_head()
            else:  # the SAME modules registered under every head{i}_* name (lfd_head.py:74-77)
                paths = tuple(getattr(self, 'head0_%s_path' % n) for n in ('classification', 'regression', 'merge'))
            for name, path in zip(('classification', 'regression', 'merge'), paths):
                setattr(self, 'head%d_%s_path' % (i, name), path)
        self._init_weights()

    def _tower(self):
        layers = []
        for i in range(self._num_conv_layers):
            cin = self._num_input_channels if i == 0 else self._num_head_channels
            k = self._conv_kernel_size
            layers.append(nn.Conv2d(cin, self._num_head_channels, kernel_size=k, stride=1, padding=k // 2, bias=self._norm_cfg is None))
            if self._norm_cfg is not None:
                layers.append(make_norm(self._norm_cfg, self._num_head_channels))
            layers.append(make_activation(self._activation_cfg))
        return layers

    def _build_head(self):
        cls_path, reg_path, merge_path = [], [], []
        if self._merge_path_flag:
            merge_path = self._tower()
        else:
            cls_path, reg_path = self._tower(), self._tower()
        c_out = self._num_classes + 1 if self._classification_loss_type == 'CrossEntropyLoss' else self._num_classes
        cls_path.append(nn.Conv2d(self._num_head_channels, c_out, kernel_size=1, stride=1, padding=0, bias=True))
        reg_path.append(nn.Conv2d(self._num_head_channels, 4, kernel_size=1, stride=1, padding=0, bias=True))
        return nn.Sequential(*cls_path), nn.Sequential(*reg_path), nn.Sequential(*merge_path)

    def _init_weights(self):  # lfd_head.py:151-162
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, mean=0, std=0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    @property
    def num_cls_channels(self):
        return self._num_classes + 1 if self._classification_loss_type == 'CrossEntropyLoss' else self._num_classes

    @property
    def uses_scale(self):
        return hasattr(self, '_scales')

    def level_paths(self, i):
        """-> (cls_tower, reg_tower, final_cls_conv, final_reg_conv); a tower is [(conv, norm)] and, with
        merge_path_flag, cls_tower is reg_tower (one shared trunk)."""
        step = 3 if self._norm_cfg is not None else 2

        def pairs(seq):
            mods = list(seq)
            return [(mods[j], mods[j + 1] if self._norm_cfg is not None else None)
                    for j in range(0, self._num_conv_layers * step, step)]
        cls_seq = getattr(self, 'head%d_classification_path' % i)
        reg_seq = getattr(self, 'head%d_regression_path' % i)
        if self._merge_path_flag:
            trunk = pairs(getattr(self, 'head%d_merge_path' % i))
            return trunk, trunk, cls_seq[0], reg_seq[0]
        return pairs(cls_seq), pairs(reg_seq), cls_seq[len(cls_seq) - 1], reg_seq[len(reg_seq) - 1]

    def forward(self, inputs):
        raise RuntimeError('LFDHead is a parameter container in lfd_b200; run it through lfd.model.LFD')
