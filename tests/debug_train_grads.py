# -*- coding: utf-8 -*-
"""Debugging aid: per-parameter gradient error of the native training step against the ATen checker (fp32 and bf16-emulated).
    python tests/debug_train_grads.py [CONFIG] [N H W]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'lfd-a-light-and-fast-detector_b200')]
import torch  # noqa: E402

import synth  # noqa: E402
from aten_train_reference import train_forward as aten_forward  # noqa: E402
from helpers import rel_err, synth_model  # noqa: E402


def main():
    linear = 'linear' in sys.argv          # fixed upstream gradients instead of the detection loss: isolates the backward arithmetic
    argv = [a for a in sys.argv if a != 'linear']
    cfg = argv[1] if len(argv) > 1 else 'WIDERFACE_XS'
    n, h, w = (int(v) for v in argv[2:5]) if len(argv) > 4 else (4, 192, 256)
    models = [synth_model(cfg, cls_bias=-2.0)[0].cuda().train() for _ in range(3)]
    x = synth.synth_input(n, h, w).cuda()
    ann = synth.synth_annotations(n, h, w, models[0]._num_classes, seed=3)
    outs = []
    for i, m in enumerate(models):
        out = m(x) if i == 0 else aten_forward(m, x, emulate_bf16=(i == 2))
        outs.append((out[0].detach().clone(), out[1].detach().clone()))
        for p in m.parameters():
            if i:
                p.grad = None
        if linear:
            if i == 0:
                gen = torch.Generator(device='cuda').manual_seed(5)
                Gc = torch.randn(out[0].shape, device='cuda', generator=gen) / out[0].numel() ** 0.5
                Gr = torch.randn(out[1].shape, device='cuda', generator=gen) / out[1].numel() ** 0.5
            ((out[0] * Gc).sum() + (out[1] * Gr).sum()).backward()
        else:
            ld = m.get_loss(out, ann)
            ld['loss'].backward()
            print('model %d loss %s' % (i, ld['loss_values']))
    for i in (1, 2):
        print('forward vs %s: cls rms %.2e reg rms %.2e' % ('fp32' if i == 1 else 'bf16-emulated', rel_err(outs[0][0], outs[i][0])[1], rel_err(outs[0][1], outs[i][1])[1]))
    print('%-55s %10s | %9s %8s | %9s %8s | %9s' % ('parameter', '|ref|', 'err fp32', 'cos', 'err emu', 'cos', 'emu-fp32'))
    for (name, p), (_, q), (_, r) in zip(*[m.named_parameters() for m in models]):
        g = p.grad.double().reshape(-1)
        row = []
        for ref in (q, r):
            t = ref.grad.double().reshape(-1)
            e = float((g - t).norm() / t.norm().clamp(min=1e-30))
            c = float((g * t).sum() / (g.norm() * t.norm()).clamp(min=1e-30))
            row += [e, c]
        ee = float((q.grad.double() - r.grad.double()).norm() / q.grad.double().norm().clamp(min=1e-30))
        print('%-55s %10.3e | %9.2e %8.5f | %9.2e %8.5f | %9.2e' % (name, float(q.grad.norm()), row[0], row[1], row[2], row[3], ee))


if __name__ == '__main__':
    main()
