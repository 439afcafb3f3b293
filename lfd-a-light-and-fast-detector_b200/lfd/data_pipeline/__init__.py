# -*- coding: utf-8 -*-
"""Only the two pieces of the reference data pipeline that touch the hot path: `Sample` and the
`simple_normalize` contract.  Dataset packing, samplers, loaders and albumentations pipelines are out of scope."""
from .dataset import Sample
from .augmentation import simple_normalize_pipeline
