"""Experiment: forward plans of several batches in flight on separate streams (throughput vs one batch at a time)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'lfd-a-light-and-fast-detector_b200'), os.path.join(ROOT, 'tests')]
import torch
from helpers import synth_model
from lfd._engine import InferencePlan

cfg = sys.argv[1] if len(sys.argv) > 1 else 'WIDERFACE_S'
N, H, W = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (8, 720, 1280)))
dev = torch.device('cuda', 0)
model, _ = synth_model(cfg)
model.to(dev).eval()
g = torch.Generator().manual_seed(1)
pool = [torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).to(dev) for _ in range(6)]
first = InferencePlan(model, N, H, W, dev, model.conv_impl, act_dtype=model.act_dtype)
caps = first.autotune() if os.environ.get('TUNE', '1') != '0' else {}
print('side-branch CTA bounds', caps)
del first
for lanes in (1, 2, 3):
    plans = [InferencePlan(model, N, H, W, dev, model.conv_impl, act_dtype=model.act_dtype) for _ in range(lanes)]
    for pl in plans:
        if caps:
            pl.apply_side_ctas(caps)
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    with torch.no_grad():
        for r in range(2):
            for i in range(len(pool)):
                for l in range(lanes):
                    with torch.cuda.stream(streams[l]):
                        plans[l].forward(pool[i], use_graph=True, slot=0)
        torch.cuda.synchronize()
        steps = 300
        t0 = time.perf_counter()
        for i in range(steps):
            l = i % lanes
            with torch.cuda.stream(streams[l]):
                plans[l].forward(pool[i % len(pool)], use_graph=True, slot=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print('%s lanes=%d: %.3f ms per batch, %.0f img/s' % (cfg, lanes, dt / steps * 1e3, N * steps / dt))
    del plans
