#!/usr/bin/env python
"""Times the tcgen05 GEMM (fo_debug_gemm, CUDA events inside the library) over tile plans for the shapes of
the streaming step; prints one line per (shape, plan).  Development aid for the cost model in fo_gemm_tc.cu."""
import itertools
import json
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    cfg = load_path_config("tiny")
    eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=2)
    shapes = [(256, 3072, 1024), (256, 1024, 1024), (256, 4096, 1024), (256, 1024, 4096), (256, 1024, 19456),
              (128, 3584, 2048), (4, 3072, 1024), (4, 1024, 4096), (512, 4096, 1024), (4096, 4096, 1024), (5120, 1024, 9216)]
    if quick:
        shapes = shapes[:2]
    if "--layer" in sys.argv:          # the four layer GEMMs of a 64-session streaming step
        shapes = shapes[:4]
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in shapes:
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        rows = []
        plans = [(-1, -1, -1), (-1, -1, -1)] + list(itertools.product((0, 1), (16, 32, 64, 128, 144, 256), (1, 2, 4, 8)))
        for (swap, bn, split) in plans:
            if swap == 1 and bn > ((M + 15) // 16) * 16 and bn != 16:
                continue
            if split > K // 64 or (M > 1024 and (split > 1 or bn < 64)):
                continue
            eng.set_option("tc_swap", swap)
            eng.set_option("tc_bn", bn)
            eng.set_option("tc_split", split)
            try:
                _, ms = eng.debug_gemm(A, W, None, backend=1, iters=30)
            except Exception as ex:
                print("fail", (M, N, K), (swap, bn, split), str(ex)[:100])
                continue
            rows.append((ms * 1e3, swap, bn, split))
        auto = min(rows[0], rows[1])
        print("auto runs", rows[0], rows[1])
        rows.sort()
        print(json.dumps({"shape": [M, N, K], "auto_us": round(auto[0], 2), "best": [[round(r[0], 2), r[1], r[2], r[3]] for r in rows[:6]],
                          "tflops_best": round(2.0 * M * N * K / rows[0][0] / 1e6, 1),
                          "weight_gbs_best": round(2.0 * N * K / rows[0][0] / 1e3, 1)}))
        sys.stdout.flush()
    eng.close()


if __name__ == "__main__":
    main()
