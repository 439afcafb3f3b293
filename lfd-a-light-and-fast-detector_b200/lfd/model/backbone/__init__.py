# -*- coding: utf-8 -*-
from .lfd_resnet import *
