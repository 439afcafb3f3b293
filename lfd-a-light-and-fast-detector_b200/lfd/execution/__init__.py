# -*- coding: utf-8 -*-
from .executor import Executor
from .utils import *
from .parallel import *
