# -*- coding: utf-8 -*-
from .focal_loss import *
from .iou_loss import *
from .cross_entropy_loss import *
from .pointwise_losses import *
