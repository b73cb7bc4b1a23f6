#!/usr/bin/env python
"""FO_TC_TRACE=1 python tools/gemm_trace.py: prints the in-kernel %globaltimer timeline of CTA (0,0,0) of the
tcgen05 GEMM for a few shapes / tile plans / producer counts (development aid).
t1 setup done, t4 first producer done issuing, t5 first stage landed, t7 last MMA issued, t8 accumulator complete,
t9 staging tile complete, t10 write-out done (ns from kernel entry)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FO_TC_TRACE"] = "1"      # needs a library built with FO_TC_TRACE_BUILD=1 python -m freeze_omni_b200.build --force
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402

cfg = load_path_config("tiny")
eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=2)
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(256, 3072, 1024), (4, 3072, 1024), (256, 1024, 4096)]:
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    for (npa, npb) in [(1, 1), (2, 1), (4, 1), (4, 2)]:
        eng.set_option("tc_npa", npa)
        eng.set_option("tc_npb", npb)
        sys.stderr.write("npa=%d npb=%d\n" % (npa, npb))
        for rep in range(2):
            eng.debug_gemm(A, W, None, backend=1, iters=0)
eng.close()
