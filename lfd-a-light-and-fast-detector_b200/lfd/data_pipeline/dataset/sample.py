# -*- coding: utf-8 -*-
__all__ = ['Sample', 'reserved_keys']

reserved_keys = ['image_bytes', 'image_type', 'image_path', 'image', 'bboxes', 'bbox_labels']


class Sample(dict):
    """dict holding one sample; reserved keys as in lfd/data_pipeline/dataset/sample.py:6-18."""

    def __str__(self):
        return 'The sample includes the following keys: \n' + ''.join('[%s]\t' % k for k in self.keys())
