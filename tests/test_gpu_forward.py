# -*- coding: utf-8 -*-
"""GPU parity of the whole forward (Gate B: against the bf16-emulated oracle, 1e-3; Gate C: drift against the
reference's fp32 outputs, reported/bounded) and of the post-process (identical kept indices)."""
import numpy as np
import pytest
import torch

import synth
from helpers import load_golden, synth_model, rel_err
from lfd import _native as nat
from oracle import lfd_oracle as orc

pytestmark = pytest.mark.gpu
FWD = ['WIDERFACE_XS', 'WIDERFACE_S', 'WIDERFACE_L', 'TT100K_L']
TOL_B = 1e-3   # BASELINE.json: outputs within 1e-3 relative (evaluated against the bf16-emulated oracle)


def _run(name, impl, graph):
    g = load_golden('forward_%s.pt' % name)
    model, sd = synth_model(name, cls_bias=g['cls_bias'], seed=g['seed'])
    model.cuda()
    model.conv_impl, model.use_cuda_graph = impl, graph
    x = synth.synth_input(g['N'], g['H'], g['W'])
    with torch.no_grad():
        cls, reg = model(x.cuda())
        if graph:  # second call replays the captured graph
            cls2, reg2 = model(x.cuda())
            assert torch.equal(cls, cls2) and torch.equal(reg, reg2)
    torch.cuda.synchronize()
    return g, sd, x, model, cls.cpu(), reg.cpu()


@pytest.mark.parametrize('impl', [nat.CONV_SIMT, nat.CONV_UMMA], ids=['simt', 'umma'])
@pytest.mark.parametrize('name', FWD)
def test_forward_matches_bf16_emulated_oracle(name, impl):
    g, sd, x, model, cls, reg = _run(name, impl, False)
    ocls, oreg, sizes = orc.forward(orc.CONFIGS[name], sd, x, emulate_bf16=True)
    assert [tuple(s) for s in sizes] == [tuple(model.head_indexes_to_feature_map_sizes[i]) for i in range(len(sizes))]
    ec, er = rel_err(cls, ocls), rel_err(reg, oreg)
    print('gate B %s: cls max/rms %.2e/%.2e reg %.2e/%.2e' % (name, ec[0], ec[1], er[0], er[1]))
    assert ec[0] < TOL_B and er[0] < TOL_B, (ec, er)
    # Gate C (reported): drift against the reference's own fp32 forward
    dc, dr = rel_err(cls, g['cls']), rel_err(reg, g['reg'])
    print('gate C %s: cls rms %.2e reg rms %.2e' % (name, dc[1], dr[1]))
    assert dc[1] < 5e-2 and dr[1] < 5e-2


@pytest.mark.parametrize('name', ['WIDERFACE_S'])
def test_forward_cuda_graph_and_u8_input(name):
    g, sd, x, model, cls, reg = _run(name, nat.CONV_UMMA, True)
    ocls, oreg, _ = orc.forward(orc.CONFIGS[name], sd, x, emulate_bf16=True)
    assert rel_err(cls, ocls)[0] < TOL_B and rel_err(reg, oreg)[0] < TOL_B
    # uint8 BGR input with the normalisation fused into the stem kernel == normalised float input
    img = np.stack([synth.synth_image_u8(g['H'], g['W'], seed=s) for s in range(g['N'])])
    with torch.no_grad():
        c8, r8 = model(torch.from_numpy(img).cuda())
        xf = torch.from_numpy(np.stack([orc.normalize_image_u8(i) for i in img])).permute(0, 3, 1, 2).contiguous()
        cf, rf = model(xf.cuda())
    assert rel_err(c8.cpu(), cf.cpu())[0] < 1e-6 and rel_err(r8.cpu(), rf.cpu())[0] < 1e-6


@pytest.mark.parametrize('name', FWD)
def test_postprocess_kept_indices_match_oracle(name):
    """Same (cls, reg) into lfd_postprocess and into the oracle's get_results: identical kept (point, class) indices
    in the same order, boxes / scores to fp32 round-off."""
    g = load_golden('forward_%s.pt' % name)
    model, _ = synth_model(name, cls_bias=g['cls_bias'], seed=g['seed'])
    model.cuda()
    model.max_detections_per_image = 32768
    for i, hw in enumerate(g['sizes']):
        model._head_indexes_to_feature_map_sizes[i] = tuple(hw)
    cls, reg = g['cls'].cuda(), g['reg'].cuda()
    cfg = orc.CONFIGS[name]
    for (thr, iou), ref in g['results'].items():
        dets, labels, src, count, overflow = model.detect((cls, reg), [m['resized_height'] for m in g['meta']],
                                                          [m['resized_width'] for m in g['meta']],
                                                          [m['resize_scale'] for m in g['meta']], thr, iou)
        assert int(overflow.item()) == 0
        _, osrc = orc.get_results(cfg, g['cls'], g['reg'], g['sizes'], g['meta'], thr, iou)
        model._classification_threshold, model._nms_cfg = thr, dict(type='nms', iou_thr=iou)
        rows = model.get_results((cls, reg), g['meta'])
        for i in range(g['N']):
            k = int(count[i].item())
            assert src[i, :k].cpu().tolist() == osrc[i].tolist(), (name, thr, iou, i)
            a, b = np.asarray(rows[i], np.float64).reshape(-1, 6), ref[i].double().numpy()   # vs the REFERENCE's get_results
            assert a.shape == b.shape
            if a.size:
                assert np.array_equal(a[:, 0], b[:, 0])
                np.testing.assert_allclose(a[:, 1:], b[:, 1:], rtol=2e-5, atol=2e-4)


def test_predict_for_single_image_runs_end_to_end():
    model, sd = synth_model('WIDERFACE_S', cls_bias=-1.0)
    img = synth.synth_image_u8(184, 248, seed=3)
    rows = model.predict_for_single_image(img, None, classification_threshold=0.2, nms_threshold=0.4)
    x = torch.from_numpy(orc.normalize_image_u8(img)).permute(2, 0, 1)[None].contiguous()
    ocls, oreg, sizes = orc.forward(orc.CONFIGS['WIDERFACE_S'], sd, x, emulate_bf16=True)
    ref, _ = orc.get_results(orc.CONFIGS['WIDERFACE_S'], ocls, oreg, sizes, [dict(resized_height=184, resized_width=248, resize_scale=1.0)], 0.2, 0.4)
    assert len(rows) > 0 and abs(len(rows) - len(ref[0])) <= max(2, len(ref[0]) // 50)
    from lfd.data_pipeline import simple_normalize_pipeline
    rows2 = model.predict_for_single_image(img, simple_normalize_pipeline, classification_threshold=0.2, nms_threshold=0.4)
    assert len(rows2) == len(rows)
