# -*- coding: utf-8 -*-
"""Launched by torchrun with 2 ranks (tests/test_gpu_train.py): three training iterations through lfd.execution's hook on
a sharded batch must leave every rank with the parameters a single process obtains on the full batch.

BatchNorm uses per-replica statistics in the reference (plain DataParallel), which makes 2 x (n/2) differ from 1 x n by
construction; the comparison therefore runs with the BatchNorm layers on their running statistics (eval mode inside the training
step, the `norm_eval` arithmetic), so that what is compared is exactly the data-parallel arithmetic of the NATIVE training step:
global positive count, per-rank loss sums, SUM all-reduce of the flat gradient buffer, fused clip + SGD.  Activations and their
gradients are bf16, so '2 x half batch' and '1 x full batch' agree to bf16 accumulation-order noise, not bit for bit.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'lfd-a-light-and-fast-detector_b200')]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import synth  # noqa: E402
from helpers import synth_model  # noqa: E402
from lfd.execution.hooks import OptimizerHook  # noqa: E402
from lfd.execution.optim import FusedSGD  # noqa: E402
from lfd.execution.parallel import shard_batch  # noqa: E402


class _Exec(object):
    def __init__(self, cfg):
        self.config_dict = cfg


def run(model, batches, sharded):
    opt = FusedSGD.from_torch(torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4), model)
    hook = OptimizerHook(dict(max_norm=10, norm_type=2, duration=5), 10)
    cfg = dict(model=model, optimizer=opt, epoch=0)
    losses = []
    for x, ann in batches:
        if sharded:
            x, ann, _ = shard_batch((x, ann, [None] * len(ann)))
        out = model(x.cuda())
        ld = model.get_loss(out, ann)
        cfg['loss'] = ld['loss']
        hook.after_train_iter(_Exec(cfg))
        losses.append(ld['loss_values']['loss'])
    return losses


def main():
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    n, h, w = 4, 192, 192
    batches = [(synth.synth_input(n, h, w, seed=10 + i), synth.synth_annotations(n, h, w, 1, seed=20 + i)) for i in range(3)]

    def fresh():
        m, _ = synth_model('WIDERFACE_XS', cls_bias=-2.0)
        m.cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eval()
        return m

    # single-process result on the full batch, computed by every rank BEFORE the process group exists
    ref = fresh()
    ref_losses = run(ref, batches, sharded=False)
    dist.init_process_group('nccl')
    ddp = fresh()
    losses = run(ddp, batches, sharded=True)
    assert ddp.loss_globally_normalised
    # LFD.get_loss logs the GLOBAL batch's loss on every rank (per-rank sums over the global positive count, all-reduced):
    # the mean over ranks of what each rank logged is that global loss, and it must equal the full-batch loss
    t = torch.tensor(losses, dtype=torch.float64, device='cuda')
    dist.all_reduce(t)
    t /= dist.get_world_size()
    worst = 0.0
    for (name, a), (_, b) in zip(ref.named_parameters(), ddp.named_parameters()):
        d = float((a.detach() - b.detach()).abs().max() / a.detach().abs().max().clamp(min=1e-12))
        worst = max(worst, d)
    ok = worst < 5e-3 and all(abs(float(t[i]) - ref_losses[i]) < 2e-3 * abs(ref_losses[i]) for i in range(len(ref_losses)))
    # every rank holds the same parameters
    flat = torch.cat([p.detach().reshape(-1) for p in ddp.parameters()])
    other = flat.clone()
    dist.broadcast(other, src=0)
    same = bool((flat == other).all())
    flag = torch.tensor([1 if (ok and same) else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('losses full batch %s | logged by the ranks (global) %s | worst relative parameter difference %.2e | identical across ranks %s'
              % (['%.5f' % v for v in ref_losses], ['%.5f' % float(v) for v in t], worst, same))
        print('DDP_OK' if int(flag.item()) == 1 else 'DDP_MISMATCH')
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == '__main__':
    sys.exit(main())
