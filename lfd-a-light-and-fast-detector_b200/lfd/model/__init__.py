# -*- coding: utf-8 -*-
from .lfd import LFD
