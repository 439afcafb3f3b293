"""Timing aid: where the end-to-end streaming loop (lfd/pipeline.py::StreamingDetector) spends its step -- forward duration and the gap
between consecutive forwards on the device, from events recorded around every forward."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'lfd-a-light-and-fast-detector_b200'), os.path.join(ROOT, 'tests')]
import torch
from helpers import synth_model
from lfd.pipeline import StreamingDetector, bind_host_to_gpu_numa_node

dev = torch.device('cuda', 0)
bind_host_to_gpu_numa_node(dev)
N, H, W = 8, 720, 1280
model, _ = synth_model('WIDERFACE_S')
model.to(dev).eval()
model.use_cuda_graph = True
model.max_detections_per_image = 8192
g = torch.Generator().manual_seed(1)
host = [torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
depth = int(os.environ.get('DEPTH', '3'))
det = StreamingDetector(model, N, H, W, 0.476, 0.3, max_out=1024, device=dev, depth=depth, copy_streams=int(os.environ.get('COPY_STREAMS', '2')))
if os.environ.get('NO_H2D'):          # diagnostic: leave the input copy out (the slots keep their first frames)
    for i in range(det.depth):
        det.stage_input(i, host[0])
    torch.cuda.synchronize()
    def _no_copy(slot, frames):
        s_ = det.slots[slot]
        with torch.cuda.stream(det.copy_stream):
            s_['h2d'].record(det.copy_stream)
    det.stage_input = _no_copy
pipe = det.pipe
orig = pipe.plan.forward
marks = []


def forward(x, use_graph=True, slot=0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(x, use_graph=use_graph, slot=slot)
    e1.record()
    marks.append((e0, e1))
    return out


with torch.no_grad():
    for i in range(12):
        det.infer(host[i % 2])
    torch.cuda.synchronize()
    for mode in (('plain',) if os.environ.get('PLAIN_ONLY') else ('plain', 'marked')):
        if mode == 'marked':
            pipe.plan.forward = forward
        marks.clear()
        steps = 300
        t0 = time.perf_counter()
        pending = []
        for i in range(steps):
            pending.append(det.submit(host[i % 2]))
            if len(pending) >= det.depth:
                det.collect(pending.pop(0))
        while pending:
            det.collect(pending.pop(0))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('%s depth %d: %.4f ms per step, %.0f img/s' % (mode, depth, dt / steps * 1e3, N * steps / dt))
    if not marks:
        sys.exit(0)
    durs = [a.elapsed_time(b) for a, b in marks[20:]]
    gaps = [marks[i][1].elapsed_time(marks[i + 1][0]) for i in range(20, len(marks) - 1)]
    print('forward on the device: mean %.4f ms (min %.4f max %.4f); gap to the next forward: mean %.4f ms (min %.4f max %.4f)'
          % (sum(durs) / len(durs), min(durs), max(durs), sum(gaps) / len(gaps), min(gaps), max(gaps)))
