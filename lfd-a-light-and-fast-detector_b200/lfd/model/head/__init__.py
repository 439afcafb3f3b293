# -*- coding: utf-8 -*-
from .lfd_head import *
