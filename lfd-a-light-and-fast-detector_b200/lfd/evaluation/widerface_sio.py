# -*- coding: utf-8 -*-
"""WIDER FACE "SIO" result writer -- WIDERFACE_train/evaluation.py:8-45 of the reference: walks the validation image tree, runs
`predict_for_single_image` on every JPEG and writes one text file per image in the format the official WIDER FACE evaluation tools read
(name, count, then `x y w h score` lines with floor / ceil rounding, preceded by one dummy box).

The reference passes `simple_widerface_val_pipeline` (normalisation on the host); here `aug_pipeline=None` sends the uint8 BGR image to the
device as is and the normalisation happens inside the stem kernel (identical arithmetic, lfd/model/lfd.py predict_for_single_image)."""
import math
import os

__all__ = ['SIO_evaluation', 'write_sio_file']


def write_sio_file(path, image_stem, results):
    """results: rows [label, score, x, y, w, h] (LFD.predict_for_single_image)."""
    with open(path, 'w') as fout:
        fout.write(image_stem + '\n')
        fout.write(str(len(results) + 1) + '\n')
        fout.write('0 0 0 0 0.001\n')
        for bbox in results:
            fout.write('%d %d %d %d %.03f' % (math.floor(bbox[2]), math.floor(bbox[3]), math.ceil(bbox[4]), math.ceil(bbox[5]),
                                              bbox[1] if bbox[1] <= 1 else 1) + '\n')


def SIO_evaluation(model, val_image_root, results_save_root='.', classification_threshold=0.5, nms_threshold=0.3, aug_pipeline=None,
                   cuda_device_index=0, verbose=True):
    assert os.path.exists(val_image_root)
    os.makedirs(results_save_root, exist_ok=True)
    counter = 0
    for parent, _, file_names in os.walk(val_image_root):
        for file_name in sorted(file_names):
            if not file_name.lower().endswith(('.jpg', '.jpeg')):
                continue
            results = model.predict_for_single_image(image=os.path.join(parent, file_name), aug_pipeline=aug_pipeline,
                                                     classification_threshold=classification_threshold, nms_threshold=nms_threshold,
                                                     class_agnostic=True, cuda_device_index=cuda_device_index)
            event_name = parent.split('/')[-1]
            os.makedirs(os.path.join(results_save_root, event_name), exist_ok=True)
            stem = file_name.split('.')[0]
            write_sio_file(os.path.join(results_save_root, event_name, stem + '.txt'), stem, results)
            counter += 1
            if verbose:
                print('[%5d] %s is processed.' % (counter, file_name))
    return counter
