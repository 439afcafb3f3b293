# -*- coding: utf-8 -*-
"""CPU tests of the host-side training / sharding plumbing (lfd.execution): hook ordering, checkpoint format, lr warm-up,
and the world_size-2 gloo path of the flat-bucket gradient all-reduce and of batch sharding."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import synth_model
from lfd.execution import parallel
from lfd.execution.hooks import Hook, LrSchedulerHook, get_priority
from lfd.execution.utils import AverageMeter, load_checkpoint, save_checkpoint


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_checkpoint_roundtrip_matches_reference_format():
    model, sd = synth_model('WIDERFACE_XS')
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[5, 7], gamma=0.1)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'w', 'epoch_3.pth')
        save_checkpoint(model, path, optimizer=opt, lr_scheduler=sched, meta=dict(epoch=3, train_iter=30))
        ck = torch.load(path, weights_only=False)
        assert set(ck) == {'meta', 'state_dict', 'optimizer_state_dict', 'lr_scheduler_state_dict'}
        assert list(ck['state_dict']) == list(sd) and 'time' in ck['meta']
        other, _ = synth_model('WIDERFACE_XS', seed=1)
        got = load_checkpoint(other, path, strict=True)
        assert got['meta']['epoch'] == 3
        for k, v in other.state_dict().items():
            assert torch.equal(v, sd[k])
        # DataParallel-style 'module.' prefixes of reference checkpoints are stripped
        torch.save(dict(meta={}, state_dict={'module.' + k: v for k, v in sd.items()}), path)
        load_checkpoint(other, path, strict=True)


def test_lr_warmup_then_scheduler():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=0.1)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2], gamma=0.1)

    class Ex(object):
        config_dict = dict(optimizer=opt, lr_scheduler=sched, train_iter=0, epoch=0)
    ex = Ex()
    hook = LrSchedulerHook(by_epoch=False, warmup_mode='linear', warmup_loops=4, warmup_ratio=0.1)
    hook.before_run(ex)
    lrs = []
    for it in range(6):
        hook.before_train_iter(ex)
        lrs.append(opt.param_groups[0]['lr'])
        ex.config_dict['train_iter'] += 1
    assert lrs[0] == pytest.approx(0.1 * (1 - 0.75 * 0.9)) and lrs[3] == pytest.approx(0.1) and lrs[5] == pytest.approx(0.1)
    assert lrs == sorted(lrs)
    assert get_priority('HIGH') < get_priority('LOW') and callable(Hook().before_run)
    m = AverageMeter()
    m.update('loss', 2.0, 3)
    m.update('loss', 4.0, 1)
    assert m.average('loss') == pytest.approx(2.5)


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    params = list(net.parameters())
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    if rank == 1:
        params[-1].grad = None   # ragged: a gradient missing on one rank still reduces with the same bucket layout
    n = parallel.allreduce_gradients(net.parameters())
    grads = [p.grad.clone() for p in net.parameters()]
    batch = (np.arange(10).reshape(5, 2), list(range(5)), [dict(i=i) for i in range(5)])
    shard = parallel.shard_batch(batch)
    total = parallel.allreduce_scalar(len(shard[1]))
    torch.save(dict(n=n, grads=grads, shard=shard[1], total=total), out % rank)
    dist.destroy_process_group()


def test_gloo_world2_flat_bucket_allreduce_and_sharding():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'r%d.pt')
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    assert r0['n'] == r1['n'] == 4 * 3 + 3 + 3 * 2 + 2
    for i, (a, b) in enumerate(zip(r0['grads'], r1['grads'])):
        assert torch.equal(a, b)
        expect = (1 + 2) * (i + 1) / 2.0 if i < 3 else 1 * (i + 1) / 2.0     # last grad existed on rank 0 only
        assert torch.allclose(a, torch.full_like(a, expect))
    assert r0['shard'] == [0, 1, 2] and r1['shard'] == [3, 4] and r0['total'] == r1['total'] == 5


def _worker_hook(rank, world_size, port, out):
    """OptimizerHook on two gloo ranks: globally normalised losses -> SUM all-reduce; per-rank means -> average."""
    from lfd.execution.hooks import OptimizerHook
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    res = {}
    for flag in (True, False):
        torch.manual_seed(0)
        net = torch.nn.Linear(3, 1, bias=False)
        net.loss_globally_normalised = flag
        opt = torch.optim.SGD(net.parameters(), lr=1.0)
        x = torch.full((1, 3), float(rank + 1))
        loss = net(x).sum()                       # d loss / d w = x  -> rank 0: 1, rank 1: 2

        class Ex(object):
            config_dict = dict(model=net, optimizer=opt, loss=loss, epoch=0)
        w0 = net.weight.detach().clone()
        OptimizerHook(None, 10).after_train_iter(Ex())
        res[flag] = (w0 - net.weight.detach()).clone()      # = the gradient that was applied (lr 1)
    torch.save(res, out % rank)
    dist.destroy_process_group()


def test_gloo_world2_optimizer_hook_sums_globally_normalised_losses():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'h%d.pt')
        port = 31500 + os.getpid() % 2000
        mp.spawn(_worker_hook, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    for flag, expect in ((True, 3.0), (False, 1.5)):       # 1 + 2 summed / averaged
        assert torch.equal(r0[flag], r1[flag])
        assert torch.allclose(r0[flag], torch.full_like(r0[flag], expect))


def test_hooks_are_registered_like_the_reference():
    """executor.py:67-99 of the reference: checkpoint LOWEST, logger VERY_LOW, lr scheduler / optimizer / evaluation NORMAL (in that
    registration order), speed LOW."""
    from lfd.execution.executor import Executor
    ex = Executor.__new__(Executor)
    ex.config_dict = dict(training_epochs=3)
    ex._hooks = []
    ex._register_all_hooks()
    assert [type(h).__name__ for h in ex._hooks] == ['LrSchedulerHook', 'OptimizerHook', 'EvaluationHook', 'SpeedHook', 'LoggerHook', 'CheckpointHook']


def _worker_sync(rank, world_size, port, out):
    """broadcast_module_state, the empty-shard step and the EvaluationHook gather on two gloo ranks."""
    from lfd.execution.hooks import EvaluationHook, OptimizerHook
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT parameters and buffers
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    net[1].running_mean.add_(float(rank))
    before = [t.clone() for t in list(net.parameters()) + list(net.buffers())]
    n = parallel.broadcast_module_state(net)
    state = [t.clone() for t in list(net.parameters()) + list(net.buffers())]
    # empty shard on rank 1: loss None, zero gradients join the SUM all-reduce
    opt = torch.optim.SGD(net.parameters(), lr=1.0)
    net.loss_globally_normalised = True
    loss = net[0](torch.ones(1, 3, 2, 2)).sum() if rank == 0 else None

    class Ex(object):
        config_dict = dict(model=net, optimizer=opt, loss=loss, epoch=0)
    w0 = net[0].weight.detach().clone()
    OptimizerHook(None, 10).after_train_iter(Ex())
    step = (w0 - net[0].weight.detach()).clone()

    class Ev(object):
        def __init__(self):
            self.seen, self.done = [], 0

        def update(self, results):
            self.seen.append(results)

        def evaluate(self):
            self.done += 1
    ev = Ev()

    class Ex2(object):
        config_dict = dict(evaluator=ev, eval_results=([[[0, 0.5 + rank, 1, 2, 3, 4]]], [dict(image_id=10 + rank)]))
    hook = EvaluationHook()
    hook.after_val_iter(Ex2())
    hook.after_val_epoch(Ex2())
    torch.save(dict(n=n, before=before, state=state, step=step, seen=ev.seen, done=ev.done), out % rank)
    dist.destroy_process_group()


def test_gloo_world2_broadcast_empty_shard_and_evaluation_gather():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 's%d.pt')
        port = 33500 + os.getpid() % 2000
        mp.spawn(_worker_sync, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    assert r0['n'] == r1['n'] > 0
    assert not all(torch.equal(a, b) for a, b in zip(r0['before'], r1['before']))
    for a, b, c in zip(r0['state'], r1['state'], r0['before']):
        assert torch.equal(a, b) and torch.equal(a, c)        # everybody holds rank 0's values
    assert torch.equal(r0['step'], r1['step']) and float(r0['step'].abs().sum()) > 0   # rank 0's gradient, summed with zeros
    for r in (r0, r1):
        (results, meta), = r['seen']
        assert [m['image_id'] for m in meta] == [10, 11] and [res[0][1] for res in results] == [0.5, 1.5] and r['done'] == 1


def test_loss_values_are_a_lazy_mapping_and_the_meter_folds_them_on_read():
    """get_loss returns `loss_values` as a Mapping whose three numbers reach the host asynchronously (lfd/model/lfd.py::LossValues; on the
    CPU they are materialised at once); AverageMeter.update_all keeps such mappings pending until an average is asked for."""
    import torch
    from lfd.model.lfd import LossValues
    from lfd.execution.utils import AverageMeter
    lv = LossValues(torch.tensor([3.0, 1.0, 2.0]))
    assert list(lv) == ['loss', 'classification_loss', 'regression_loss'] and len(lv) == 3
    assert lv['loss'] == 3.0 and dict(lv.items()) == dict(loss=3.0, classification_loss=1.0, regression_loss=2.0)
    assert 'missing' not in lv and lv.get('missing', 7) == 7
    with pytest.raises(KeyError):
        lv['missing']
    m = AverageMeter()
    m.update_all(lv, 2)
    m.update_all(LossValues(torch.tensor([6.0, 2.0, 4.0])), 1)
    assert m._pending and not m._sum                       # nothing has been read yet
    assert abs(m.average('loss') - (3.0 * 2 + 6.0) / 3) < 1e-12
    assert abs(m.averages()['regression_loss'] - (2.0 * 2 + 4.0) / 3) < 1e-12
    m.update('loss', 9.0, 1)                               # the reference's eager update still works
    assert abs(m.average('loss') - (3.0 * 2 + 6.0 + 9.0) / 4) < 1e-12
    m.reset()
    assert not m._pending and not m.averages()
