# -*- coding: utf-8 -*-
"""LFD options outside the BASELINE configs on the GPU against vectors produced by the REFERENCE's own modules
(tests/gen_golden_variants.py -> tests/golden/loss_variants.pt): GIoU / DIoU / CIoU, SmoothL1 / MSE on 'independent' targets,
BCE-with-logits, quality focal loss, distance_to_bbox_mode 'exp', range_assign_mode 'shorter' -- losses, their gradients w.r.t. the network
outputs and the decoded + NMS'd results -- plus the stand-alone box-loss modules, FastBlock / FastestBlock backbones and head variants."""
import numpy as np
import pytest
import torch

import synth
from helpers import build_model, load_golden, rel_err
from lfd.model import losses as L
from oracle import lfd_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = None


def gold():
    global GOLD
    if GOLD is None:
        GOLD = load_golden('loss_variants.pt')
    return GOLD


def make_loss(name):
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


VARIANT_NAMES = ['focal_iou_exp', 'focal_giou_sigmoid', 'focal_diou_exp', 'focal_ciou_sigmoid', 'bce_iou_sigmoid', 'qfl_giou_sigmoid',
                 'focal_smoothl1_independent', 'ce_mse_independent', 'focal_ciou_exp_shorter']


@pytest.mark.parametrize('name', VARIANT_NAMES)
def test_loss_variant_matches_reference(name):
    g = gold()[name]
    v = g['variant']
    model = build_model(v['cfg'])
    model._classification_loss_func = make_loss(v['cls'])
    model._regression_loss_func = make_loss(v['reg'])
    model._regression_loss_type = 'independent' if v['reg'] in ('SmoothL1Loss', 'MSELoss') else 'union'
    model._distance_to_bbox_mode = v['bbox']
    model._range_assign_mode = v.get('assign', model._range_assign_mode)
    model.cuda().eval()
    for i, s in enumerate(g['sizes']):
        model._head_indexes_to_feature_map_sizes[i] = tuple(s)
    cls = g['cls_pred'].cuda().requires_grad_(True)
    reg = g['reg_pred'].cuda().requires_grad_(True)
    ld = model.get_loss((cls, reg), g['ann'])
    ld['loss'].backward()
    for k, want in g['loss_values'].items():
        assert abs(ld['loss_values'][k] - want) <= 3e-5 * abs(want) + 1e-6, (name, k, ld['loss_values'][k], want)
    ec, er = rel_err(cls.grad.cpu(), g['grad_cls']), rel_err(reg.grad.cpu(), g['grad_reg'])
    assert ec[0] < 2e-4 and er[0] < 5e-4, (name, ec, er)
    # decode + class-aware NMS (lfd.py:434-509): same kept rows as the reference's get_results
    r = g['results']
    model._classification_threshold, model._nms_cfg = r['thr'], dict(type='nms', iou_thr=r['iou'])
    meta = [dict(resized_height=g['H'], resized_width=g['W'], resize_scale=1.0) for _ in range(g['N'])]
    model.max_detections_per_image = 8192
    rows = model.get_results((cls.detach(), reg.detach()), meta)
    for got, want in zip(rows, r['rows']):
        assert len(got) == len(want), (name, len(got), len(want))
        if not want:
            continue
        a, b = np.asarray(got, np.float64), np.asarray(want, np.float64)
        assert np.array_equal(a[:, 0], b[:, 0]), name                     # labels, in kept (score-descending) order
        assert np.allclose(a[:, 1:], b[:, 1:], rtol=2e-5, atol=2e-3), (name, float(np.abs(a[:, 1:] - b[:, 1:]).max()))


@pytest.mark.parametrize('name', ['IoULoss', 'GIoULoss', 'DIoULoss', 'CIoULoss'])
def test_standalone_box_loss_modules(name):
    b = gold()['box_pairs']
    pred = b['pred'].cuda().requires_grad_(True)
    mod = getattr(L, name)(eps=1e-6, reduction='mean', loss_weight=1.0)
    el = mod(pred, b['target'].cuda(), reduction_override='none')
    el.sum().backward()
    assert float((el.detach().cpu() - b[name]['loss']).abs().max()) < 2e-5 * max(1.0, float(b[name]['loss'].abs().max()))
    assert rel_err(pred.grad.cpu(), b[name]['grad'])[0] < 2e-4
    # reductions of the module (mean with avg_factor, loss_weight)
    mod2 = getattr(L, name)(eps=1e-6, reduction='mean', loss_weight=2.0)
    val = mod2(b['pred'].cuda(), b['target'].cuda(), avg_factor=7.0)
    assert abs(float(val) - 2.0 * float(b[name]['loss'].sum()) / 7.0) < 1e-4 * abs(float(val))


def test_pointwise_modules_match_torch():
    torch.manual_seed(3)
    x, t = torch.randn(50, 4).cuda(), torch.randn(50, 4).cuda()
    assert torch.allclose(L.SmoothL1Loss(beta=0.5)(x, t), torch.nn.functional.smooth_l1_loss(x, t, beta=0.5))
    assert torch.allclose(L.MSELoss()(x, t), torch.nn.functional.mse_loss(x, t))
    lab = torch.randint(0, 5, (50,)).cuda()          # 4 classes, label 4 = background
    tgt = torch.nn.functional.one_hot(lab, 5)[:, :4].float()
    assert torch.allclose(L.BCEWithLogitsLoss()(x, lab), torch.nn.functional.binary_cross_entropy_with_logits(x, tgt))
