#!/usr/bin/env python
"""Period of a dependent chain of trivially small kernels inside a CUDA graph on this GPU: the floor that ~175
kernels per streaming step cannot go below (development aid; uses torch only to launch the tiny kernels)."""
import torch

x = torch.zeros(32, device="cuda")
big = torch.zeros(256, 1024, device="cuda")
for name, fn in (("add_ on 32 floats", lambda: x.add_(1.0)), ("add_ on 256x1024 floats", lambda: big.add_(1.0))):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(400):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.2f us per kernel in a 400-kernel graph chain" % (name, e0.elapsed_time(e1) * 1e3 / 4000))
