# -*- coding: utf-8 -*-
"""SimpleNeck -- constructor / state_dict compatible with lfd/model/neck/simple_neck.py:18-74."""
import torch.nn as nn

from ..backbone.lfd_resnet import make_norm, make_activation, init_conv_norm

__all__ = ['SimpleNeck']


class SimpleNeck(nn.Module):

    def __init__(self, num_neck_channels, num_input_channels_list, num_input_strides_list,
                 norm_cfg=dict(type='BatchNorm2d'), activation_cfg=dict(type='ReLU', inplace=True)):
        super(SimpleNeck, self).__init__()
        assert len(num_input_channels_list) == len(num_input_strides_list)
        self._num_neck_channels = num_neck_channels
        self._num_input_channels_list = num_input_channels_list
        self._num_input_strides_list = num_input_strides_list
        self._norm_cfg, self._activation_cfg = norm_cfg, activation_cfg
        self._num_inputs = len(num_input_channels_list)
        for i, ch in enumerate(num_input_channels_list):
            layers = [nn.Conv2d(ch, num_neck_channels, kernel_size=1, stride=1, padding=0, bias=norm_cfg is None)]
            if norm_cfg is not None:
                layers.append(make_norm(norm_cfg, num_neck_channels))
            layers.append(make_activation(activation_cfg))
            setattr(self, 'neck%d' % i, nn.Sequential(*layers))
        init_conv_norm(self)

    @property
    def num_output_strides_list(self):
        return self._num_input_strides_list

    def level(self, i):
        seq = getattr(self, 'neck%d' % i)
        return seq[0], (seq[1] if self._norm_cfg is not None else None)

    def forward(self, inputs):
        raise RuntimeError('SimpleNeck is a parameter container in lfd_b200; run it through lfd.model.LFD')
