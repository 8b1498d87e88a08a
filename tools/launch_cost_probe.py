#!/usr/bin/env python
"""Host cost of one asynchronous CUDA call on this box (kernel launch, event record, memset) -- the per-batch overhead of the pipeline is
~25 such calls per batch and k.  Prints microseconds per call."""
import time

import torch

x = torch.zeros(1024, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2000)]
torch.cuda.synchronize()
for name, fn in (("kernel launch (x.add_)", lambda i: x.add_(1.0)), ("event record", lambda i: ev[i].record()), ("memset (x.zero_)", lambda i: x.zero_())):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(2000):
            fn(i)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
    print(f"{name}: {(t1 - t0) / 2000 * 1e6:.1f} us per call (host side, queue not full: 2000 calls)")
