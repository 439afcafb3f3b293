# -*- coding: utf-8 -*-
import numpy

__all__ = ['simple_normalize_pipeline']


def simple_normalize_pipeline(sample):
    """(x/255 - 0.5)/0.5 on the BGR uint8 image -> float32 HWC
    (lfd/data_pipeline/augmentation/augmentation_pipeline.py:31-36, albumentations.Normalize(mean=.5, std=.5))."""
    img = sample['image'].astype(numpy.float32)
    sample['image'] = (img - numpy.float32(127.5)) * numpy.float32(1.0 / 127.5)
    return sample
