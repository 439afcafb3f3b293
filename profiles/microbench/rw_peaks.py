"""HBM streams by access mix on this GPU (CUDA events, best of 10): write-only (fill_), read-only (sum), copy (read + write).
The layer rooflines in bench.py divide by the driver's copy figure (MEASURED_PEAKS.json).  Measured: fp32 fill_ 6.9 TB/s, fp32 sum 5.5-5.8 TB/s,
copy 6.5 TB/s -- a layer that mostly writes (the stem: 22 MB in, 236 MB out) or mostly reads has NO lower ceiling than the copy figure.
torch's bf16 fill_ kernel tops out at 3.6-3.9 TB/s; that is the kernel (2-byte elements), not the memory system."""
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


print('write only (bf16 fill_ 1 GiB) %.0f GB/s   <- torch kernel limit, not HBM' % best(lambda: a.fill_(1.0), 2 * n))
print('write only (fp32 fill_ 1 GiB) %.0f GB/s' % best(lambda: a.view(torch.float32).fill_(1.0), 2 * n))
print('read only  (fp32 sum 1 GiB)   %.0f GB/s' % best(lambda: a.view(torch.float32).sum(), 2 * n))
print('copy       (read + write)     %.0f GB/s' % best(lambda: b.copy_(a), 4 * n))
small = torch.empty(118 * 1024 * 1024, dtype=torch.bfloat16, device=dev)     # 236 MB: the stem output of WIDERFACE-S 720p batch 8
print('write only (fp32 fill_ 236 MB) %.0f GB/s' % best(lambda: small.view(torch.float32).fill_(1.0), small.numel() * 2))
