# coding: utf-8
from .nms import *
