#!/usr/bin/env python
"""Development check of the persistent stack kernel (option stack_kernel=1) against the per-kernel chain on the same
inputs: two sets of sessions fed identical PCM, one per path; prints the max-abs differences and the step times."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = load_path_config("shipped")
eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=2 * S,
             max_stream_frames=cfg.chunk_feat_frames)
eng.set_option("use_graph", graph)
a = eng.alloc(S)
b = eng.alloc(S)
g = torch.Generator().manual_seed(7)
worst = 0.0
for i in range(steps):
    pcm = (0.05 * torch.randn(S, cfg.samples_per_chunk, generator=g) * 32768).round().to(torch.int16).cuda()
    eng.set_option("stack_kernel", 1)
    n0 = eng.get_option("stack_launches")
    e1, y1 = eng.stream_step(a, pcm, 1.0)
    torch.cuda.synchronize()
    assert eng.get_option("stack_launches") == n0 + 1, "stack kernel did not launch"
    eng.set_option("stack_kernel", 0)
    e0, y0 = eng.stream_step(b, pcm, 1.0)
    torch.cuda.synchronize()
    de = (e1.float() - e0.float()).abs().max().item()
    dy = (y1.float() - y0.float()).abs().max().item()
    worst = max(worst, de, dy)
    if i < 3 or i == steps - 1 or de > 1e-2:
        print("step %d: enc diff %.3e (max %.2f)  adapter diff %.3e" % (i, de, e0.abs().max().item(), dy), flush=True)
print("worst %.3e" % worst)
pcm = (0.05 * torch.randn(S, cfg.samples_per_chunk, generator=g) * 32768).round().to(torch.int16).cuda()
for flag in (0, 1, 0, 1):
    eng.set_option("stack_kernel", flag)
    eng.set_option("use_graph", 1)
    for _ in range(5):
        eng.stream_step(a, pcm, 1.0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.stream_step(a, pcm, 1.0)
    e1.record()
    torch.cuda.synchronize()
    print("stack_kernel=%d: %.3f ms/step" % (flag, e0.elapsed_time(e1) / 50), flush=True)
eng.close()
