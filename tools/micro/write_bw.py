"""Pure-write bandwidth of the B200 (torch fill_) at sizes around the PPO-update activation store (193 MB)."""
import torch
dev = torch.device("cuda:0")
for mb in (48, 96, 193, 400, 1000, 4000):
    x = torch.empty(mb * 1000 * 1000 // 4, dtype=torch.float32, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x.fill_(1.0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    y = torch.empty_like(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_(); e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    tc = e0.elapsed_time(e1)
    print("%5d MB: fill %.1f us = %.2f TB/s ; copy %.1f us = %.2f TB/s (read+write)" % (mb, 1e3 * t, mb * 1e6 / t / 1e9, 1e3 * tc, 2 * mb * 1e6 / tc / 1e9))
