#!/usr/bin/env python
"""Times the fat offline GEMM shapes on the persistent and the one-tile-per-CTA tcgen05 kernels (fo_debug_gemm, graph-timed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402

cfg = load_path_config("tiny")
eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=2)
g = torch.Generator().manual_seed(1)
for (M, N, K) in [(23936, 4096, 1024), (23936, 1024, 4096), (23936, 3072, 1024), (23936, 1024, 1024)]:
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    row = []
    for persist in (1, 0):
        eng.set_option("tc_persist", persist)
        _, ms = eng.debug_gemm(A, W, b, backend=1, iters=20)
        row.append("%s %.1f us %.0f TF" % ("persist" if persist else "tile/CTA", ms * 1e3, 2.0 * M * N * K / ms / 1e9))
    print((M, N, K), " | ".join(row), flush=True)
eng.close()
