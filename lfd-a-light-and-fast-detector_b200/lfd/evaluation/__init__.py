# -*- coding: utf-8 -*-
from .base_evaluator import *
from .coco_evaluator import *
from .widerface_sio import *
