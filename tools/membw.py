#!/usr/bin/env python
"""Device memory bandwidth by access type (torch kernels, CUDA events): the write-heavy propagate kernel's
roofline context.  read: sum; write: fill_; copy: copy_ (read+write bytes)."""
import torch

def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for mb in (1258, 4096):
    n = mb * (1 << 20) // 4
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    a.normal_()
    w = t(lambda: b.fill_(1.0)); r = t(lambda: a.sum()); c = t(lambda: b.copy_(a))
    print(f"{mb} MiB: write {n*4/w/1e6:.0f} GB/s ({w*1e3:.0f} us)  read {n*4/r/1e6:.0f} GB/s  copy {2*n*4/c/1e6:.0f} GB/s (r+w)")
    # write-heavy mix like propagate+collide: read 1 part, write 3 parts
    k = n // 4
    m = t(lambda: (b.fill_(1.0), a[:k].sum()))
    print(f"   fill {mb} MiB then read {mb//4} MiB back to back: {(n+k)*4/m/1e6:.0f} GB/s")
