# -*- coding: utf-8 -*-
"""Hook bus of the reference executor (lfd/execution/hooks/*.py) in one module: priorities, optimizer step with gradient
clipping (optimizer_hook.py:26-36) -- with the data-parallel gradient all-reduce inserted between backward and clip --,
lr warm-up + scheduler stepping (lr_scheduler_hook.py:55-100), speed, logging and checkpoint hooks."""
import time

import torch
from torch.nn.utils import clip_grad

from .optim import FusedSGD
from .parallel import allreduce_gradients, world

PRIORITIES = dict(HIGHEST=0, VERY_HIGH=10, HIGH=30, NORMAL=50, LOW=70, VERY_LOW=90, LOWEST=100)


def get_priority(priority):
    if isinstance(priority, int):
        if not 0 <= priority <= 100:
            raise ValueError('priority must be between 0 and 100')
        return priority
    if isinstance(priority, str):
        return PRIORITIES[priority.upper()]
    raise TypeError('priority must be an integer or a priority name')


class Hook(object):
    priority = PRIORITIES['NORMAL']
    STAGES = ('before_run', 'after_run', 'before_train_epoch', 'after_train_epoch', 'before_val_epoch', 'after_val_epoch',
              'before_train_iter', 'after_train_iter', 'before_val_iter', 'after_val_iter')

    def __getattr__(self, name):
        if name in Hook.STAGES:
            return lambda executor: None
        raise AttributeError(name)


class OptimizerHook(Hook):
    def __init__(self, grad_clip_cfg, training_epochs):
        assert isinstance(grad_clip_cfg, dict) or grad_clip_cfg is None
        self._grad_clip_cfg = dict(grad_clip_cfg) if grad_clip_cfg is not None else None
        if self._grad_clip_cfg is not None:
            self._grad_clip_duration = self._grad_clip_cfg.pop('duration', training_epochs)
            assert isinstance(self._grad_clip_duration, int) and self._grad_clip_duration > 0

    def after_train_iter(self, executor):
        cfg = executor.config_dict
        model, opt = cfg['model'], cfg['optimizer']
        opt.zero_grad()
        if cfg['loss'] is not None:        # None: this rank's shard of the batch was empty; it still joins the all-reduce with zeros
            cfg['loss'].backward()
        # one flat-bucket all-reduce (no-op for a single process); losses already normalised by the global number of
        # positives (LFD.get_loss) add up over ranks, otherwise the per-rank means are averaged
        average = not getattr(model, 'loss_globally_normalised', False)
        clip = self._grad_clip_cfg is not None and cfg['epoch'] < self._grad_clip_duration
        if isinstance(opt, FusedSGD):
            # native path: the parameters / gradients are views into flat buffers (lfd/_train.py) -- the bucket IS the gradient
            # storage, the averaging factor, the clip and the SGD update are one fused pass (lfd_grad_sqnorm + lfd_sgd_step)
            flat = opt._sync()
            _, w = world()
            if w > 1:
                torch.distributed.all_reduce(flat.grad, op=torch.distributed.ReduceOp.SUM)
            if clip and (self._grad_clip_cfg.get('norm_type', 2) not in (2, 2.0)):
                raise NotImplementedError('FusedSGD clips with the 2-norm (every shipped config)')
            norm = opt.step(max_norm=float(self._grad_clip_cfg['max_norm']) if clip else 0.0, grad_scale=(1.0 / w) if (average and w > 1) else 1.0)
            cfg['grad_norm'] = norm if norm is not None else 0
        else:
            allreduce_gradients(model.parameters(), average=average)
            if self._grad_clip_cfg is not None:
                if clip:
                    params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
                    cfg['grad_norm'] = clip_grad.clip_grad_norm_(params, **self._grad_clip_cfg) if params else 0
                else:
                    cfg['grad_norm'] = 0
            opt.step()
        if hasattr(model, 'invalidate_plans'):
            model.invalidate_plans()


class LrSchedulerHook(Hook):
    def __init__(self, by_epoch=True, warmup_mode=None, warmup_loops=0, warmup_ratio=0.1):
        assert warmup_mode in (None, 'constant', 'linear', 'exp')
        self._by_epoch, self._warmup_mode, self._warmup_loops, self._warmup_ratio = by_epoch, warmup_mode, warmup_loops, warmup_ratio
        self._skips = 0

    def before_run(self, executor):
        for g in executor.config_dict['optimizer'].param_groups:
            g.setdefault('initial_lr', g['lr'])
        self._base_lr = [g['initial_lr'] for g in executor.config_dict['optimizer'].param_groups]

    def get_warmup_lr(self, loop):
        if self._warmup_mode == 'constant':
            return [lr * self._warmup_ratio for lr in self._base_lr]
        if self._warmup_mode == 'linear':
            k = (1 - loop / self._warmup_loops) * (1 - self._warmup_ratio)
            return [lr * (1 - k) for lr in self._base_lr]
        k = self._warmup_ratio ** (1 - loop / self._warmup_loops)
        return [lr * k for lr in self._base_lr]

    def _tick(self, executor, loop):
        opt = executor.config_dict['optimizer']
        if self._warmup_mode is not None and loop <= self._warmup_loops:
            for g, lr in zip(opt.param_groups, self.get_warmup_lr(loop)):
                g['lr'] = lr
        elif self._skips >= 0 and loop == self._warmup_loops + 1 and self._warmup_mode is not None:
            for g, lr in zip(opt.param_groups, self._base_lr):
                g['lr'] = lr
            for _ in range(self._skips):
                executor.config_dict['lr_scheduler'].step()
            self._skips = 0

    def before_train_epoch(self, executor):
        if self._by_epoch:
            self._tick(executor, executor.config_dict['epoch'] + 1)

    def before_train_iter(self, executor):
        if not self._by_epoch:
            self._tick(executor, executor.config_dict['train_iter'] + 1)

    def after_train_epoch(self, executor):
        loop = executor.config_dict['train_iter'] if not self._by_epoch else executor.config_dict['epoch']
        if self._warmup_mode is not None and loop <= self._warmup_loops:
            self._skips += 1           # the scheduler does not step while warming up
        else:
            executor.config_dict['lr_scheduler'].step()


class SpeedHook(Hook):
    def before_train_iter(self, executor):
        self._t = time.time()

    def after_train_iter(self, executor):
        dt = max(time.time() - self._t, 1e-9)
        executor.config_dict['speed'] = executor.config_dict['batch_size'] * world()[1] / dt


def _fmt_norm(v):
    return '%.4f' % float(v) if torch.is_tensor(v) else v


class LoggerHook(Hook):
    def after_train_iter(self, executor):
        cfg = executor.config_dict
        if cfg['train_iter'] % max(cfg.get('display_interval', 100), 1) == 0:
            avg = cfg['train_average_meter'].averages()
            cfg['logger'].info('epoch %d iter %d lr %.6f speed %.1f img/s grad_norm %s %s' % (
                cfg['epoch'], cfg['train_iter'], executor.get_current_lr(), cfg.get('speed', 0.0), _fmt_norm(cfg.get('grad_norm', '-')),
                ' '.join('%s %.5f' % kv for kv in avg.items())))
            cfg['train_average_meter'].reset()

    def after_val_epoch(self, executor):
        cfg = executor.config_dict
        cfg['logger'].info('val epoch %d: %s' % (cfg['epoch'], ' '.join('%s %.5f' % kv for kv in cfg['val_average_meter'].averages().items())))
        cfg['val_average_meter'].reset()


class EvaluationHook(Hook):
    """reference lfd/execution/hooks/evaluation_hook.py: feed every validation batch's results to the evaluator, evaluate at
    the end of the validation epoch.  One process per GPU: the per-rank result shards are all-gathered (python objects, small)
    so that every rank's evaluator sees the whole batch in rank order."""

    def after_val_iter(self, executor):
        cfg = executor.config_dict
        if cfg.get('evaluator') is None:
            return
        results, meta = cfg['eval_results']
        rank, ws = world()
        if ws > 1:
            import torch.distributed as dist
            shards = [None] * ws
            dist.all_gather_object(shards, (results, meta))
            results = [r for sh in shards for r in sh[0]]
            meta = [m for sh in shards for m in sh[1]]
        cfg['evaluator'].update((results, meta))

    def after_val_epoch(self, executor):
        if executor.config_dict.get('evaluator') is not None:
            executor.config_dict['evaluator'].evaluate()


class CheckpointHook(Hook):
    def after_train_epoch(self, executor):
        cfg = executor.config_dict
        if cfg.get('save_interval', 0) > 0 and cfg['epoch'] % cfg['save_interval'] == 0 and world()[0] == 0:
            executor.save()
