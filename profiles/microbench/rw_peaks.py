"""HBM ceilings by access mix on this GPU (CUDA events, best of 10): write-only (fill_), read-only (sum), copy (read + write).
The layer rooflines in bench.py divide by the driver's copy figure (MEASURED_PEAKS.json); a layer that mostly WRITES (the stem: 22 MB in,
236 MB out) or mostly READS is bounded by the one-directional figure printed here."""
import torch

dev = torch.device('cuda', 0)
n = 1 << 29     # 1 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)


def best(fn, bytes_moved, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_moved / (min(t) * 1e-3) / 1e9


print('write only (fill_ 1 GiB)      %.0f GB/s' % best(lambda: a.fill_(1.0), 2 * n))
print('read only  (sum 1 GiB)        %.0f GB/s' % best(lambda: a.view(torch.int16).sum(), 2 * n))
print('copy       (read + write)     %.0f GB/s' % best(lambda: b.copy_(a), 4 * n))
small = torch.empty(118 * 1024 * 1024, dtype=torch.bfloat16, device=dev)     # 236 MB: the stem output of WIDERFACE-S 720p batch 8
print('write only (fill_ 236 MB)     %.0f GB/s' % best(lambda: small.fill_(1.0), small.numel() * 2))
