# -*- coding: utf-8 -*-
"""Golden vectors for the LFD options outside the BASELINE configs, produced by the REFERENCE's own
modules (imported from /root/reference on the CPU, same stubs as tests/gen_golden.py):

    LFD.get_loss (lfd/model/lfd.py:284-395) + autograd with  GIoULoss / DIoULoss / CIoULoss (losses/iou_loss.py:125-283,324-430),
    SmoothL1Loss / MSELoss on 'independent' targets (losses/smooth_l1_loss.py, mse_loss.py; lfd.py:219-220,353-358),
    BCEWithLogitsLoss (losses/bce_with_logits_loss.py) and QualityFocalLoss (losses/gfocal_loss.py:10-50),
    distance_to_bbox_mode 'exp', range_assign_mode 'shorter';
    LFD.get_results (lfd.py:397-509) for the 'exp' and 'independent' decodes.

    python tests/gen_golden_variants.py      ->  tests/golden/loss_variants.pt   (needs /root/reference; the file is committed)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import gen_golden as gg  # noqa: E402
import synth  # noqa: E402
from oracle import lfd_oracle as orc  # noqa: E402

VARIANTS = [
    dict(name='focal_iou_exp', cfg='WIDERFACE_S', cls='FocalLoss', reg='IoULoss', bbox='exp'),
    dict(name='focal_giou_sigmoid', cfg='WIDERFACE_S', cls='FocalLoss', reg='GIoULoss', bbox='sigmoid'),
    dict(name='focal_diou_exp', cfg='WIDERFACE_S', cls='FocalLoss', reg='DIoULoss', bbox='exp'),
    dict(name='focal_ciou_sigmoid', cfg='WIDERFACE_S', cls='FocalLoss', reg='CIoULoss', bbox='sigmoid'),
    dict(name='bce_iou_sigmoid', cfg='WIDERFACE_S', cls='BCEWithLogitsLoss', reg='IoULoss', bbox='sigmoid'),
    dict(name='qfl_giou_sigmoid', cfg='WIDERFACE_S', cls='QualityFocalLoss', reg='GIoULoss', bbox='sigmoid'),
    dict(name='focal_smoothl1_independent', cfg='WIDERFACE_S', cls='FocalLoss', reg='SmoothL1Loss', bbox='sigmoid'),
    dict(name='ce_mse_independent', cfg='TT100K_S', cls='CrossEntropyLoss', reg='MSELoss', bbox='sigmoid'),
    dict(name='focal_ciou_exp_shorter', cfg='WIDERFACE_S', cls='FocalLoss', reg='CIoULoss', bbox='exp', assign='shorter'),
]


def make_loss(R, name):
    L = R['losses']
    if name == 'FocalLoss':
        return L.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0)
    if name == 'CrossEntropyLoss':
        return L.CrossEntropyLoss(reduction='mean', loss_weight=1.0)
    if name == 'BCEWithLogitsLoss':
        return L.BCEWithLogitsLoss(reduction='mean', loss_weight=1.0)
    if name == 'QualityFocalLoss':
        return L.QualityFocalLoss(use_sigmoid=True, beta=2.0, reduction='mean', loss_weight=1.0)
    if name == 'SmoothL1Loss':
        return L.SmoothL1Loss(beta=0.11, reduction='mean', loss_weight=2.0)
    if name == 'MSELoss':
        return L.MSELoss(reduction='mean', loss_weight=1.5)
    return getattr(L, name)(eps=1e-6, reduction='mean', loss_weight=1.0)


# reference forward / results / loss goldens of the configurations outside the BASELINE list: the shipped TrafficLight model (head without
# norm layers) and two small nets covering FastBlock / FastestBlock, the 'fastest' stem and 3x3 head towers
EXTRA_FORWARD_CASES = {'TL_L': (1, 136, 200, -1.0), 'TEST_FAST': (2, 120, 168, -1.0), 'TEST_FASTEST': (2, 152, 200, -1.0)}


def main():
    R = gg.import_reference()
    torch.set_num_threads(8)
    for name, (n, h, w, cls_bias) in EXTRA_FORWARD_CASES.items():
        gg.forward_case(R, name, n, h, w, cls_bias, os.path.join(HERE, 'golden'))
    out = {}
    H, W, N = 256, 320, 2
    for v in VARIANTS:
        cfg = orc.CONFIGS[v['cfg']]
        model = gg.build_ref_model(R, cfg)
        C = cfg['lfd']['num_classes']
        model._classification_loss_func = make_loss(R, v['cls'])
        model._regression_loss_func = make_loss(R, v['reg'])
        model._regression_loss_type = 'independent' if v['reg'] in ('SmoothL1Loss', 'MSELoss') else 'union'
        model._distance_to_bbox_mode = v['bbox']
        model._range_assign_mode = v.get('assign', cfg['lfd']['range_assign_mode'])
        sizes = gg.sizes_for(cfg, H, W)
        for i, s in enumerate(sizes):
            model._head_indexes_to_feature_map_sizes[i] = s
        P = sum(h * w for h, w in sizes)
        Cp = C + 1 if v['cls'] == 'CrossEntropyLoss' else C
        g = torch.Generator().manual_seed(1234)
        cls_pred = (torch.randn(N, P, Cp, generator=g) * 1.5 - 1.0).requires_grad_(True)
        if model._regression_loss_type == 'independent':
            reg_pred = (torch.randn(N, P, 4, generator=g) * 0.3 + 0.3).requires_grad_(True)
        elif v['bbox'] == 'exp':
            reg_pred = (torch.randn(N, P, 4, generator=g) * 0.6 + 2.5).requires_grad_(True)
        else:
            reg_pred = (torch.randn(N, P, 4, generator=g) * 1.0).requires_grad_(True)
        ann = synth.synth_annotations(N, H, W, C, seed=11, max_boxes=8)
        ld = model.get_loss((cls_pred, reg_pred), ann)
        ld['loss'].backward()
        meta = [dict(resized_height=H, resized_width=W, resize_scale=1.0) for _ in range(N)]
        thr, iou = 0.6, 0.4
        model._classification_threshold = thr
        model._nms_cfg = dict(type='nms', iou_thr=iou)
        with torch.no_grad():
            rows = model.get_results((cls_pred.detach(), reg_pred.detach()), meta)
        out[v['name']] = dict(variant=v, H=H, W=W, N=N, sizes=sizes, ann=ann, cls_pred=cls_pred.detach().clone(), reg_pred=reg_pred.detach().clone(),
                              loss_values={k: float(x) for k, x in ld['loss_values'].items()}, grad_cls=cls_pred.grad.clone(), grad_reg=reg_pred.grad.clone(),
                              results=dict(thr=thr, iou=iou, rows=rows))
        print('%-28s loss %s  |grad| %.3e / %.3e  detections %s' % (v['name'], ld['loss_values'], float(cls_pred.grad.norm()), float(reg_pred.grad.norm()),
                                                                     [len(r) for r in rows]))
    # element-wise box losses of the stand-alone modules (reduction='none') on random box pairs, with autograd gradients
    g = torch.Generator().manual_seed(77)
    n = 300
    c = torch.rand(n, 2, generator=g) * 200 + 20
    wh = torch.rand(n, 2, generator=g) * 80 + 4
    target = torch.cat([c - wh / 2, c + wh / 2], 1)
    pred = target + torch.randn(n, 4, generator=g) * 12
    pred[:40] = target[:40] + 300          # disjoint pairs
    pred[40:60, 2:] = pred[40:60, :2] + wh[40:60] * 0.5
    box = dict(pred=pred.clone(), target=target.clone())
    for name in ('IoULoss', 'GIoULoss', 'DIoULoss', 'CIoULoss'):
        pr = pred.clone().requires_grad_(True)
        mod = getattr(R['losses'], name)(eps=1e-6, reduction='mean', loss_weight=1.0)
        el = mod(pr, target, reduction_override='none')
        el.sum().backward()
        box[name] = dict(loss=el.detach().clone(), grad=pr.grad.clone())
    out['box_pairs'] = box
    torch.save(out, os.path.join(HERE, 'golden', 'loss_variants.pt'))


if __name__ == '__main__':
    main()
