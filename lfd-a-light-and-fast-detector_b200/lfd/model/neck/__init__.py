# -*- coding: utf-8 -*-
from .simple_neck import *
