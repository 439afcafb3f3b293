# -*- coding: utf-8 -*-
"""GPU parity of the non-conv entry points of the C-ABI: NMS, sigmoid focal loss, label assignment, detection loss."""
import numpy as np
import pytest
import torch

from helpers import load_golden, synth_model, rel_err
from oracle import lfd_oracle as orc

pytestmark = pytest.mark.gpu


def test_nms_known_answers_and_random():
    from lfd.model.utils import nms
    k = load_golden('known_answers.pt')
    _, inds = nms(torch.from_numpy(k['nms_doc_dets']).cuda(), 0.6)
    assert inds.tolist() == [0, 3, 4]                                   # nms.py:24-34
    suppressed, inds2 = nms(k['nms_doc_dets'], 0.6, device_id=0)       # numpy in -> numpy out
    assert isinstance(inds2, np.ndarray) and inds2.tolist() == [0, 3, 4] and len(suppressed) == 3
    _, inds = nms(torch.from_numpy(k['nms_rand_dets']).cuda(), k['nms_rand_thr'])
    assert inds.tolist() == k['nms_rand_keep'].tolist()
    _, inds = nms(torch.zeros((0, 5)).cuda(), 0.5)
    assert inds.numel() == 0
    rng = np.random.RandomState(5)
    for n in (1, 33, 5000, 9000):   # 5000 / 9000: global-memory sort + sweep path
        d = np.concatenate([rng.uniform(0, 400, (n, 2)), rng.uniform(2, 80, (n, 2)), rng.uniform(-1, 1, (n, 1))], 1).astype(np.float32)
        d[:, 2:4] += d[:, 0:2]
        _, inds = nms(torch.from_numpy(d).cuda(), 0.45)
        assert inds.tolist() == orc.nms(d, 0.45).tolist()


def test_multiclass_nms_matches_oracle():
    from lfd.model.utils import multiclass_nms
    rng = np.random.RandomState(9)
    n, C = 600, 5
    boxes = np.concatenate([rng.uniform(0, 300, (n, 2)), rng.uniform(4, 90, (n, 2))], 1).astype(np.float32)
    boxes[:, 2:] += boxes[:, :2]
    scores = rng.uniform(0, 1, (n, C)).astype(np.float32)
    padded = np.concatenate([scores, np.zeros((n, 1), np.float32)], 1)
    dets, labels = multiclass_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(padded).cuda(), 0.3, dict(type='nms', iou_thr=0.5))
    odets, olabels, _ = orc.multiclass_nms(boxes, scores, 0.3, 0.5)
    assert labels.tolist() == olabels.tolist()
    np.testing.assert_allclose(dets.cpu().numpy(), odets, rtol=1e-6, atol=1e-4)


def test_sigmoid_focal_loss_module():
    from lfd.model.losses import FocalLoss
    g = torch.Generator().manual_seed(2)
    for C in (1, 45):
        x = (torch.randn((777, C), generator=g) * 3)
        t = torch.randint(0, C + 1, (777,), generator=g)
        xg = x.clone().cuda().requires_grad_(True)
        loss = FocalLoss(gamma=2.0, alpha=0.25)(xg, t.cuda(), avg_factor=13.0)
        loss.backward()
        ref = orc.sigmoid_focal_loss_forward(x, t, 2.0, 0.25)
        assert abs(float(loss) - float(ref.sum() / 13.0)) < 1e-5 * abs(float(ref.sum() / 13.0))
        gref = orc.sigmoid_focal_loss_backward(x, t, torch.full_like(x, 1.0 / 13.0), 2.0, 0.25)
        assert rel_err(xg.grad.cpu(), gref)[0] < 1e-5
        # element-wise, fp32 round-off only
        el = FocalLoss(reduction='none')(x.cuda(), t.cuda())
        assert float((el.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('name', ['WIDERFACE_S', 'WIDERFACE_L', 'TT100K_L'])
def test_label_assignment_matches_reference_golden(name):
    """lfd_assign_targets against the REFERENCE's annotation_to_target output: structure exact, green scores to
    2 ulp (torch's CPU sqrt is not correctly rounded), regression targets of positive rows exact."""
    g = load_golden('assign_%s.pt' % name)
    model, _ = synth_model(name)
    model.cuda()
    pts = model.generate_point_coordinates(dict(enumerate(g['sizes'])))
    ct, rt = model.annotation_to_target(pts, [torch.from_numpy(a[0]) for a in g['ann']], [torch.from_numpy(a[1]) for a in g['ann']])
    ct, rt = ct.cpu().numpy(), rt.cpu().numpy()
    cfg = orc.CONFIGS[name]
    for i, ((boxes, labels), im) in enumerate(zip(g['ann'], g['images'])):
        nz = np.nonzero(np.abs(ct[i]).sum(-1) > 0)[0]
        assert np.array_equal(nz, im['nz_rows'].numpy())
        ref = im['nz_cls'].numpy()
        assert np.array_equal(ct[i][nz] == -1, ref == -1) and np.array_equal(np.sign(ct[i][nz]), np.sign(ref))
        np.testing.assert_allclose(ct[i][nz], ref, rtol=3e-7, atol=0)
        pos = im['pos_rows'].numpy()
        assert np.array_equal(rt[i][pos], im['pos_reg'].numpy())
        oct_, ort = orc.assign_targets(cfg, g['sizes'], boxes, labels)   # and bit-exact against the IEEE oracle
        assert np.array_equal(ct[i], oct_) and np.array_equal(rt[i], ort)


@pytest.mark.parametrize('name', ['WIDERFACE_XS', 'WIDERFACE_S', 'WIDERFACE_L', 'TT100K_L'])
def test_get_loss_matches_reference_golden(name):
    g = load_golden('forward_%s.pt' % name)
    model, _ = synth_model(name, cls_bias=g['cls_bias'])
    model.cuda()
    for i, hw in enumerate(g['sizes']):
        model._head_indexes_to_feature_map_sizes[i] = tuple(hw)
    cls = g['cls'].clone().cuda().requires_grad_(True)
    reg = g['reg'].clone().cuda().requires_grad_(True)
    out = model.get_loss((cls, reg), g['ann'])
    out['loss'].backward()
    lv, mv = g['loss_values'], out['loss_values']
    for key in ('loss', 'classification_loss', 'regression_loss'):
        assert abs(mv[key] - lv[key]) <= 2e-5 * max(abs(lv[key]), 1e-3), (key, mv[key], lv[key])
    assert rel_err(cls.grad.cpu(), g['grad_cls'])[0] < 2e-5
    if float(g['grad_reg'].abs().max()) > 0:
        assert rel_err(reg.grad.cpu(), g['grad_reg'])[0] < 1e-4
    else:
        assert float(reg.grad.abs().max()) == 0.0


def test_batched_nms_matches_oracle():
    """batched_nms (utils/nms.py:119-158): the label * (max coordinate + 1) offsets live inside the native NMS kernel; kept indices and
    rows against the oracle's NMS on the offset boxes."""
    from lfd.model.utils import batched_nms
    rng = np.random.RandomState(21)
    n = 300
    xy = rng.uniform(0, 120, (n, 2)).astype(np.float32)
    wh = rng.uniform(8, 50, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)
    scores = rng.uniform(0.05, 1.0, n).astype(np.float32)
    labels = rng.randint(0, 4, n).astype(np.int64)
    dets, keep = batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), dict(type='nms', iou_thr=0.4))
    off = labels.astype(np.float32) * np.float32(boxes.max() + np.float32(1))
    okeep = orc.nms(np.concatenate([boxes + off[:, None], scores[:, None]], 1).astype(np.float32), 0.4)
    assert keep.cpu().tolist() == list(map(int, okeep))
    # like the reference (nms.py:152-156) the returned boxes are (box + offset) - offset in fp32, not the input boxes bit for bit
    shifted = (boxes + off[:, None]).astype(np.float32)
    want = (shifted[okeep] - off[okeep][:, None]).astype(np.float32)
    assert np.array_equal(dets[:, :4].cpu().numpy(), want) and np.array_equal(dets[:, 4].cpu().numpy(), scores[okeep])
    # class agnostic: plain NMS
    dets2, keep2 = batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(),
                               dict(type='nms', iou_thr=0.4, class_agnostic=True))
    assert keep2.cpu().tolist() == list(map(int, orc.nms(np.concatenate([boxes, scores[:, None]], 1).astype(np.float32), 0.4)))


@pytest.mark.gpu
@pytest.mark.parametrize('n,n_cls,skew', [(3000, 45, False), (6000, 12, True), (1500, 2, True)])
def test_batched_nms_many_candidates_several_classes(n, n_cls, skew):
    """More than 1024 candidates with several classes take the per-class sweeps of nms_kernel (one warp per class, whole CTA for classes
    with > 512 candidates): same kept set, order and rows as the oracle's single NMS over the class-offset boxes (nms.py:141-156)."""
    from lfd.model.utils import batched_nms
    rng = np.random.RandomState(n + n_cls)
    xy = rng.uniform(0, 900, (n, 2)).astype(np.float32)
    wh = rng.uniform(10, 120, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)
    scores = rng.uniform(0.05, 1.0, n).astype(np.float32)
    scores[rng.randint(0, n, n // 20)] = np.float32(0.5)                 # score ties: resolved by the input index in both
    if skew:                                                             # one class holds most of the boxes (> 512: the CTA-wide sweep)
        labels = np.where(rng.uniform(size=n) < 0.7, 0, rng.randint(0, n_cls, n)).astype(np.int64)
    else:
        labels = rng.randint(0, n_cls, n).astype(np.int64)
    dets, keep = batched_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), dict(type='nms', iou_thr=0.45))
    off = labels.astype(np.float32) * np.float32(boxes.max() + np.float32(1))
    shifted = (boxes + off[:, None]).astype(np.float32)
    okeep = orc.nms(np.concatenate([shifted, scores[:, None]], 1).astype(np.float32), 0.45)
    assert keep.cpu().tolist() == list(map(int, okeep))
    want = (shifted[okeep] - off[okeep][:, None]).astype(np.float32)
    assert np.array_equal(dets[:, :4].cpu().numpy(), want) and np.array_equal(dets[:, 4].cpu().numpy(), scores[okeep])
