"""Timing aid: forward plan alone, post-process alone, both pipelined (bench path) for one workload."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'lfd-a-light-and-fast-detector_b200'), os.path.join(ROOT, 'tests')]
import torch
from helpers import synth_model
from lfd.pipeline import ForwardPostPipeline

cfg = sys.argv[1] if len(sys.argv) > 1 else 'TT100K_L'
N, H, W = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (16, 1080, 1920)))
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 2e-4
dev = torch.device('cuda', 0)
model, _ = synth_model(cfg)
model.to(dev).eval()
model.max_detections_per_image = 8192
g = torch.Generator().manual_seed(1)
pool = [torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).to(dev) for _ in range(4)]
plan = model.inference_plan(N, H, W, dev)
for i, hw in enumerate(plan.level_sizes):
    model._head_indexes_to_feature_map_sizes[i] = hw
post = model.post_plan(N, plan.level_sizes, dev)
post.set_meta([W] * N, [H] * N, [1.0] * N)


def timeit(fn, n=200):
    for _ in range(24):          # covers every (input buffer, output slot) pair: graph instantiation stays outside the timed loop
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    cls, reg = plan.forward(pool[0], use_graph=False)
    scores = cls.sigmoid() if plan.cls_channels == model._num_classes else cls.softmax(-1)[..., :-1]
    thr = float(torch.quantile(scores.flatten()[:2000000].float(), 1.0 - frac))
    k = [0]

    def fwd():
        k[0] += 1
        plan.forward(pool[k[0] % 4], use_graph=True, slot=0)
    print('%s forward %.3f ms' % (cfg, timeit(fwd)))
    print('%s post    %.3f ms (thr %.4f, %s kept)' % (cfg, timeit(lambda: post.run(cls, reg, thr, 0.3)), thr, post.count[:N].tolist()[:4]))
    print('%s post with no candidate passing (score / threshold scan only) %.3f ms' % (cfg, timeit(lambda: post.run(cls, reg, 2.0, 0.3))))
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                post.run(cls, reg, thr, 0.3)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=8, max_name_column_width=60))
    except Exception as e:
        print('profiler unavailable: %r' % (e,))
    pipe = ForwardPostPipeline(model, plan, post, thr, 0.3)
    model.use_cuda_graph = True

    def both():
        k[0] += 1
        pipe.enqueue(pool[k[0] % 4])
    print('%s pipelined forward + post %.3f ms per batch' % (cfg, timeit(both)))
