"""Host<->device copy bandwidth of the box (context for the e2e number): pinned 512 MB buffers, CUDA events."""
import torch
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1), "GB/s")
