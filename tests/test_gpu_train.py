# -*- coding: utf-8 -*-
"""Training step through the drop-in API on the GPU.

What is native here: label assignment, the losses and their gradients w.r.t. the network outputs, the flat-bucket gradient
all-reduce.  The conv-stack forward / backward in training mode is ATen / cuDNN (lfd/_train.py, library code) -- these tests
tie that module-graph walk to the native layer plan and check that the train loop of lfd.execution actually trains.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import synth
from helpers import rel_err, synth_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _freeze_norms(model):
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.GroupNorm)):
            m.eval()


@pytest.mark.parametrize('cfg', ['WIDERFACE_XS', 'TT100K_S'])
def test_train_graph_walk_matches_native_plan(cfg):
    """With the BatchNorm layers on running statistics the ATen walk computes the same function as the native (bf16) plan:
    same wiring, taps, shared towers, scales and output layout; the difference is the bounded bf16 drift of DESIGN.md gate C."""
    model, _ = synth_model(cfg)
    model.cuda()
    x = synth.synth_input(2, 184, 248).cuda()
    model.eval()
    with torch.no_grad():
        cls_n, reg_n = model(x)
    sizes_eval = dict(model._head_indexes_to_feature_map_sizes)
    model.train()
    _freeze_norms(model)
    with torch.no_grad():
        cls_t, reg_t = model(x)
    assert cls_t.shape == cls_n.shape and reg_t.shape == reg_n.shape
    assert dict(model._head_indexes_to_feature_map_sizes) == sizes_eval
    for a, b in ((cls_n, cls_t), (reg_n, reg_t)):
        _, rms = rel_err(a, b)
        assert rms < 2.5e-2, rms


def test_train_loop_reduces_the_loss():
    """Executor-style iterations on a fixed batch: forward (train mode, batch statistics), native get_loss, backward,
    gradient clipping, SGD step -- all parameters receive finite gradients and the loss goes down."""
    model, _ = synth_model('WIDERFACE_XS', cls_bias=-2.0)
    model.cuda().train()
    n, h, w = 4, 256, 256
    x = synth.synth_input(n, h, w).cuda()
    ann = synth.synth_annotations(n, h, w, 1, seed=3)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    losses = []
    for it in range(8):
        out = model(x)
        ld = model.get_loss(out, ann)
        opt.zero_grad()
        ld['loss'].backward()
        if it == 0:
            for name, p in model.named_parameters():
                assert p.grad is not None and torch.isfinite(p.grad).all(), name
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10, norm_type=2)
        opt.step()
        losses.append(ld['loss_values']['loss'])
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < 0.8 * losses[0], losses
    # the eval path picks the updated weights up (the plan cache is keyed by a parameter fingerprint)
    model.eval()
    with torch.no_grad():
        cls, reg = model(x)
    assert torch.isfinite(cls).all() and torch.isfinite(reg).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (NCCL)')
def test_ddp_two_ranks_match_one_rank_on_the_full_batch():
    """tests/run_train_ddp.py: 2 ranks (NCCL), each on half of the batch, global positive-count normalisation + SUM
    all-reduce of one flat gradient bucket == 1 rank on the whole batch."""
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tests', 'run_train_ddp.py')]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'DDP_OK' in r.stdout, r.stdout[-3000:]
