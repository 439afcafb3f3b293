# -*- coding: utf-8 -*-
"""COCOEvaluator -- constructor / update / evaluate / get_eval_display_str of lfd/evaluation/coco_evaluator.py:13-83, fed by
Executor.val with (LFD.get_results rows, meta_batch).

The reference delegates the metric to pycocotools (COCO / COCOeval, iouType='bbox', params.maxDets = [100, 300, 1000]); that package is a
third-party dependency which is neither vendored in the reference nor installed here (reference README: `pycocotools`, unpinned), so the
bounding-box protocol it publishes is restated in `CocoBoxEval` below: per image and category, detections sorted by score and cut at
maxDets, greedy matching in score order against the ground truths at the 10 IoU thresholds 0.50:0.05:0.95 (crowd regions match any number of
detections with intersection / detection-area overlap and are ignored, ground truths outside the area range are ignored and are matched
last), precision made monotone and sampled at 101 recall points, AP = mean over thresholds / categories of the valid entries.
Parity is anchored on the reference's call site (coco_evaluator.py:58-80: the six summary numbers mAP, mAP_50, mAP_75, mAP_s, mAP_m, mAP_l at
maxDets[2]) and on hand-computed cases in tests/test_evaluation.py (no golden from pycocotools can be produced in this image: "parity
unpinned" against the package itself).  Host-side numpy code: this is bookkeeping after the device post-process, not part of the hot path."""
import json
import os

import numpy as np

from .base_evaluator import Evaluator

__all__ = ['COCOEvaluator', 'CocoBoxEval']


def _iou_xywh(dt, gt, crowd):
    """dt [D,4], gt [G,4] (x, y, w, h) -> [D,G]; for crowd ground truths the union is the detection's area."""
    if len(dt) == 0 or len(gt) == 0:
        return np.zeros((len(dt), len(gt)))
    dx1, dy1, dx2, dy2 = dt[:, 0:1], dt[:, 1:2], dt[:, 0:1] + dt[:, 2:3], dt[:, 1:2] + dt[:, 3:4]
    gx1, gy1, gx2, gy2 = gt[:, 0], gt[:, 1], gt[:, 0] + gt[:, 2], gt[:, 1] + gt[:, 3]
    iw = np.clip(np.minimum(dx2, gx2) - np.maximum(dx1, gx1), 0, None)
    ih = np.clip(np.minimum(dy2, gy2) - np.maximum(dy1, gy1), 0, None)
    inter = iw * ih
    da, ga = dt[:, 2:3] * dt[:, 3:4], gt[:, 2] * gt[:, 3]
    union = np.where(np.asarray(crowd, bool)[None, :], da + 0 * ga[None, :], da + ga[None, :] - inter)
    return inter / np.maximum(union, 1e-12)


class CocoBoxEval(object):
    """Bounding-box AP of the COCO detection protocol."""
    IOU_THRS = np.linspace(0.5, 0.95, 10)
    REC_THRS = np.linspace(0.0, 1.0, 101)
    AREA_RNG = dict(all=(0.0, 1e10), small=(0.0, 32.0 ** 2), medium=(32.0 ** 2, 96.0 ** 2), large=(96.0 ** 2, 1e10))

    def __init__(self, gts, dts, image_ids, category_ids, max_dets=(100, 300, 1000)):
        """gts: dicts(image_id, category_id, bbox xywh, area, iscrowd); dts: dicts(image_id, category_id, bbox xywh, score)."""
        self.image_ids, self.category_ids, self.max_dets = sorted(set(image_ids)), sorted(set(category_ids)), list(max_dets)
        self._gt, self._dt = {}, {}
        for g in gts:
            self._gt.setdefault((g['image_id'], g['category_id']), []).append(g)
        for d in dts:
            self._dt.setdefault((d['image_id'], d['category_id']), []).append(d)

    def _evaluate_image(self, img, cat, area, max_det):
        gts, dts = self._gt.get((img, cat), []), self._dt.get((img, cat), [])
        if not gts and not dts:
            return None
        g_ignore = np.array([bool(g.get('iscrowd', 0)) or bool(g.get('ignore', 0)) or not (area[0] <= g['area'] <= area[1]) for g in gts], bool)
        order = np.argsort(g_ignore, kind='mergesort')                 # ignored ground truths are matched last
        gts, g_ignore = [gts[i] for i in order], g_ignore[order]
        dorder = np.argsort([-d['score'] for d in dts], kind='mergesort')[:max_det]
        dts = [dts[i] for i in dorder]
        crowd = [bool(g.get('iscrowd', 0)) for g in gts]
        ious = _iou_xywh(np.array([d['bbox'] for d in dts], np.float64).reshape(-1, 4), np.array([g['bbox'] for g in gts], np.float64).reshape(-1, 4), crowd)
        T, D, G = len(self.IOU_THRS), len(dts), len(gts)
        gtm, dtm, dt_ig = -np.ones((T, G), int), -np.ones((T, D), int), np.zeros((T, D), bool)
        for ti, t in enumerate(self.IOU_THRS):
            for di in range(D):
                best, m = min(t, 1 - 1e-10), -1
                for gi in range(G):
                    if gtm[ti, gi] >= 0 and not crowd[gi]:
                        continue
                    if m > -1 and not g_ignore[m] and g_ignore[gi]:
                        break                                            # a regular match is never traded for an ignored one
                    if ious[di, gi] < best:
                        continue
                    best, m = ious[di, gi], gi
                if m == -1:
                    continue
                dt_ig[ti, di] = g_ignore[m]
                dtm[ti, di], gtm[ti, m] = m, di
        d_area = np.array([d['bbox'][2] * d['bbox'][3] for d in dts])
        out_of_range = ~((area[0] <= d_area) & (d_area <= area[1])) if D else np.zeros((0,), bool)
        dt_ig = dt_ig | ((dtm < 0) & out_of_range[None, :])
        return dict(scores=np.array([d['score'] for d in dts]), matched=dtm >= 0, dt_ignore=dt_ig, n_gt=int((~g_ignore).sum()))

    def evaluate(self):
        """-> dict(mAP, mAP_50, mAP_75, mAP_s, mAP_m, mAP_l) at max_dets[-1]; -1 where no ground truth exists."""
        T, R, K = len(self.IOU_THRS), len(self.REC_THRS), len(self.category_ids)
        max_det = self.max_dets[-1]
        precision = {a: -np.ones((T, R, K)) for a in self.AREA_RNG}
        for a, rng in self.AREA_RNG.items():
            for k, cat in enumerate(self.category_ids):
                per_img = [e for e in (self._evaluate_image(img, cat, rng, max_det) for img in self.image_ids) if e is not None]
                if not per_img:
                    continue
                n_gt = sum(e['n_gt'] for e in per_img)
                if n_gt == 0:
                    continue
                scores = np.concatenate([e['scores'] for e in per_img])
                order = np.argsort(-scores, kind='mergesort')
                matched = np.concatenate([e['matched'] for e in per_img], 1)[:, order]
                ignore = np.concatenate([e['dt_ignore'] for e in per_img], 1)[:, order]
                tps = np.cumsum(matched & ~ignore, 1).astype(np.float64)
                fps = np.cumsum(~matched & ~ignore, 1).astype(np.float64)
                for ti in range(T):
                    tp, fp = tps[ti], fps[ti]
                    rc = tp / n_gt
                    pr = tp / np.maximum(tp + fp, np.spacing(1))
                    for i in range(len(pr) - 1, 0, -1):                 # precision envelope
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    inds = np.searchsorted(rc, self.REC_THRS, side='left')
                    q = np.zeros(R)
                    ok = inds < len(pr)
                    q[ok] = pr[inds[ok]]
                    precision[a][ti, :, k] = q

        def mean(p, ti=None):
            p = p if ti is None else p[ti:ti + 1]
            v = p[p > -1]
            return float(v.mean()) if v.size else -1.0
        return dict(mAP=mean(precision['all']), mAP_50=mean(precision['all'], 0), mAP_75=mean(precision['all'], 5),
                    mAP_s=mean(precision['small']), mAP_m=mean(precision['medium']), mAP_l=mean(precision['large']))


class COCOEvaluator(Evaluator):

    def __init__(self, annotation_path, label_indexes_to_category_ids):
        assert os.path.isfile(annotation_path), 'annotation file does not exist!!!'
        assert isinstance(label_indexes_to_category_ids, dict), 'label index to category id must be a dict!!!'
        with open(annotation_path, 'r') as f:
            self._dataset = json.load(f)
        self._label_indexes_to_category_ids = label_indexes_to_category_ids
        self._detection_results = list()
        self._image_ids = set()
        self._eval_display_str = ''
        self.stats = None

    def update(self, results):
        """results: (predict_bboxes, meta_batch); predict_bboxes[i] = rows [label, score, x, y, w, h] of image i (LFD.get_results)."""
        assert isinstance(results, tuple) and len(results) == 2, 'update info should contain two parts: predict bboxes and meta info.'
        predict_bboxes, meta_batch = results
        for rows, meta in zip(predict_bboxes, meta_batch):
            image_id = meta['image_id']
            for row in rows:
                self._image_ids.add(image_id)
                self._detection_results.append(dict(image_id=image_id, bbox=[float(v) for v in row[2:6]], score=float(row[1]),
                                                    category_id=self._label_indexes_to_category_ids[row[0]]))

    def evaluate(self):
        self._eval_display_str = '\n'
        if len(self._detection_results) == 0:
            self._eval_display_str += 'No bboxes detected! Evaluation abort!\n'
            return
        imgs = set(self._image_ids)
        gts = [dict(a, area=a.get('area', a['bbox'][2] * a['bbox'][3])) for a in self._dataset.get('annotations', []) if a['image_id'] in imgs]
        cats = [c['id'] for c in self._dataset.get('categories', [])] or sorted(set(g['category_id'] for g in gts))
        ev = CocoBoxEval(gts, self._detection_results, imgs, cats, max_dets=(100, 300, 1000))
        self.stats = ev.evaluate()
        for metric in ('mAP', 'mAP_50', 'mAP_75', 'mAP_s', 'mAP_m', 'mAP_l'):
            self._eval_display_str += '{:<10}:{:.5f}\n'.format(metric, self.stats[metric])
        self._detection_results.clear()

    def get_eval_display_str(self):
        return self._eval_display_str
