# -*- coding: utf-8 -*-
"""Debug aid: true start / end of every plan op INSIDE a CUDA-graph replay (%globaltimer stamps written by the kernels).

Needs a trace build:   LFD_B200_TIMELINE=1 LFD_B200_OUT=$PWD/build_variants/lib_timeline.so python lfd-a-light-and-fast-detector_b200/build.py --force
                       then run with LFD_B200_LIB=build_variants/lib_timeline.so (TUNE=1: after InferencePlan.autotune)
usage: python tests/debug_timeline.py [config] [batch] [H] [W]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'lfd-a-light-and-fast-detector_b200')]
import torch  # noqa: E402

from lfd import _native as nat  # noqa: E402


def main():
    from helpers import synth_model
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'WIDERFACE_S'
    n, h, w = [int(v) for v in (sys.argv[2:5] + ['8', '720', '1280'][len(sys.argv[2:5]):])]
    model, _ = synth_model(cfg)
    model.cuda()
    plan = model.inference_plan(n, h, w, torch.device('cuda', 0))
    if os.environ.get('TUNE'):
        print('autotune: side-branch CTA bounds', plan.autotune())
    rows = plan.describe()
    k = len(rows)
    buf = torch.zeros((k, 2), dtype=torch.int64, device='cuda')
    nat.lib().lfd_debug_set_timeline(nat.ptr(buf))          # before the first forward: the pointers are baked into the graph
    x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device='cuda')
    for _ in range(5):
        plan.forward(x, use_graph=True)
    torch.cuda.synchronize()
    # back-to-back replays (no host synchronisation in between: an idle GPU starts its next kernel slowly); the stamps that
    # survive are those of the last replay
    hi = torch.iinfo(torch.int64).max
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(12):
        buf[:, 0].fill_(hi)
        buf[:, 1].zero_()
        if i == 11:
            e0.record()
        plan.forward(x, use_graph=True)
    e1.record()
    torch.cuda.synchronize()
    runs = [(e0.elapsed_time(e1), buf.cpu().clone())]
    nat.lib().lfd_debug_set_timeline(None)
    runs.sort(key=lambda r: r[0])
    ms, t = runs[len(runs) // 2]
    t0 = int(t[:, 0].min())
    print('graph replay %.3f ms (events); ops %d; span of stamps %.3f ms' % (ms, k, (int(t[:, 1].max()) - t0) * 1e-6))
    order = sorted(range(k), key=lambda i: int(t[i, 0]))
    busy = 0.0
    for i in order:
        r = rows[i]
        s, e = (int(t[i, 0]) - t0) * 1e-3, (int(t[i, 1]) - t0) * 1e-3
        busy += e - s
        print('%3d  %-10s k%d s%d %3d->%3d%s %4dx%-4d res=%d  start %8.1f us  end %8.1f us  dur %6.1f us  %s' % (
            i, r['kind'], r['ksize'], r['stride'], r['Cin'], r['Cout'], ('->%3d' % r['tail_cout']) if r['tail_cout'] else ('+sc  ' if r.get('ds_cout') else '     '),
            r['Ho'], r['Wo'], int(r['res']), s, e, e - s, r.get('query') or ''))
    print('sum of durations %.1f us' % busy)


if __name__ == '__main__':
    main()
