# -*- coding: utf-8 -*-
"""lfd.evaluation: the COCO bounding-box protocol restated in CocoBoxEval (hand-computed cases: pycocotools is not installed, so
parity with the package itself is unpinned -- see the module docstring), the COCOEvaluator interface the executor drives, and
the WIDER FACE SIO text format."""
import json
import os

import pytest

from lfd.evaluation import COCOEvaluator, CocoBoxEval
from lfd.evaluation.widerface_sio import write_sio_file


def _gt(img, cat, box, crowd=0):
    return dict(image_id=img, category_id=cat, bbox=list(box), area=box[2] * box[3], iscrowd=crowd)


def _dt(img, cat, box, score):
    return dict(image_id=img, category_id=cat, bbox=list(box), score=score)


def test_perfect_and_disjoint_detections():
    gts = [_gt(1, 1, (10, 10, 50, 50)), _gt(1, 1, (100, 100, 20, 20)), _gt(2, 2, (5, 5, 200, 120))]
    dts = [_dt(g['image_id'], g['category_id'], g['bbox'], 0.9 - 0.1 * i) for i, g in enumerate(gts)]
    s = CocoBoxEval(gts, dts, [1, 2], [1, 2]).evaluate()
    assert s['mAP'] == pytest.approx(1.0) and s['mAP_50'] == pytest.approx(1.0) and s['mAP_75'] == pytest.approx(1.0)
    assert s['mAP_s'] == pytest.approx(1.0) and s['mAP_m'] == pytest.approx(1.0) and s['mAP_l'] == pytest.approx(1.0)   # 400 / 2500 / 24000 px^2
    far = [_dt(d['image_id'], d['category_id'], (d['bbox'][0] + 500, d['bbox'][1], d['bbox'][2], d['bbox'][3]), d['score']) for d in dts]
    s = CocoBoxEval(gts, far, [1, 2], [1, 2]).evaluate()
    assert s['mAP'] == 0.0 and s['mAP_50'] == 0.0


def test_hand_computed_precision_recall():
    """two ground truths; detections in score order: hit, miss, hit -> precision 1 up to recall 0.5, 2/3 up to recall 1:
    AP = (51 * 1 + 50 * 2/3) / 101 at every IoU threshold the hits clear (they are exact boxes)."""
    gts = [_gt(1, 1, (0, 0, 40, 40)), _gt(1, 1, (100, 0, 40, 40))]
    dts = [_dt(1, 1, (0, 0, 40, 40), 0.9), _dt(1, 1, (300, 300, 40, 40), 0.8), _dt(1, 1, (100, 0, 40, 40), 0.7)]
    s = CocoBoxEval(gts, dts, [1], [1]).evaluate()
    want = (51 * 1.0 + 50 * (2.0 / 3.0)) / 101
    assert s['mAP'] == pytest.approx(want, abs=1e-9) and s['mAP_50'] == pytest.approx(want, abs=1e-9)


def test_iou_thresholds_crowd_and_max_dets():
    # a detection with IoU 0.6 counts at thresholds 0.50, 0.55, 0.60 only -> mAP = 3/10, mAP_50 = 1, mAP_75 = 0
    gts = [_gt(1, 1, (0, 0, 100, 100))]
    dts = [_dt(1, 1, (0, 0, 100, 60), 0.9)]                      # intersection 6000 / union 10000
    s = CocoBoxEval(gts, dts, [1], [1]).evaluate()
    assert s['mAP_50'] == pytest.approx(1.0) and s['mAP_75'] == 0.0 and s['mAP'] == pytest.approx(0.3)
    # detections inside a crowd region are ignored (neither TP nor FP), whatever their number
    gts = [_gt(1, 1, (0, 0, 40, 40)), _gt(1, 1, (200, 200, 300, 300), crowd=1)]
    dts = [_dt(1, 1, (0, 0, 40, 40), 0.9), _dt(1, 1, (210, 210, 30, 30), 0.95), _dt(1, 1, (260, 260, 30, 30), 0.93)]
    assert CocoBoxEval(gts, dts, [1], [1]).evaluate()['mAP'] == pytest.approx(1.0)
    # only the max_dets best-scored detections of an image are looked at
    dts = [_dt(1, 1, (500 + 50 * i, 0, 40, 40), 0.99 - 0.001 * i) for i in range(5)] + [_dt(1, 1, (0, 0, 40, 40), 0.5)]
    assert CocoBoxEval(gts[:1], dts, [1], [1], max_dets=(1, 2, 5)).evaluate()['mAP'] == 0.0
    assert CocoBoxEval(gts[:1], dts, [1], [1], max_dets=(1, 2, 6)).evaluate()['mAP'] > 0.0


def test_coco_evaluator_interface(tmp_path):
    ann = dict(images=[dict(id=7), dict(id=8)], categories=[dict(id=3), dict(id=5)],
               annotations=[dict(id=1, image_id=7, category_id=3, bbox=[10, 10, 60, 60], area=3600, iscrowd=0),
                            dict(id=2, image_id=8, category_id=5, bbox=[0, 0, 30, 30], area=900, iscrowd=0)])
    path = os.path.join(str(tmp_path), 'instances.json')
    json.dump(ann, open(path, 'w'))
    ev = COCOEvaluator(path, {0: 3, 1: 5})
    ev.evaluate()
    assert 'No bboxes detected' in ev.get_eval_display_str()
    # rows as LFD.get_results returns them: [label, score, x, y, w, h]
    ev.update(([[[0, 0.9, 10, 10, 60, 60]], [[1, 0.8, 0, 0, 30, 30], [0, 0.3, 100, 100, 10, 10]]], [dict(image_id=7), dict(image_id=8)]))
    ev.evaluate()
    assert ev.stats['mAP_50'] == pytest.approx(1.0)
    text = ev.get_eval_display_str()
    assert text.count('\n') == 7 and 'mAP_50    :1.00000' in text and ev._detection_results == []


def test_sio_file_format(tmp_path):
    p = os.path.join(str(tmp_path), 'a.txt')
    write_sio_file(p, '0_Parade_1', [[0, 0.98765, 10.7, 20.2, 30.1, 40.9], [0, 1.2, 1.0, 2.0, 3.0, 4.0]])
    assert open(p).read().splitlines() == ['0_Parade_1', '3', '0 0 0 0 0.001', '10 20 31 41 0.988', '1 2 3 4 1.000']
