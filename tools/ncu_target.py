#!/usr/bin/env python
"""Target program for ncu captures (profiles/r02_*): warms the context up, then runs exactly ONE step between
cudaProfilerStart / cudaProfilerStop (use `ncu --profile-from-start off`).

    python tools/ncu_target.py stream [sessions]     one eager 64-session streaming step (181 kernels), windows full
    python tools/ncu_target.py offline [B]           one full-utterance encode of B x 30 s (fbank + encoder + adapter)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "stream"
    cfg = load_path_config("shipped")
    if mode == "stream":
        S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=S,
                     max_stream_frames=cfg.chunk_feat_frames)
        ids = eng.alloc(S)
        g = torch.Generator().manual_seed(5)
        pcm = (0.05 * torch.randn(4, S, cfg.samples_per_chunk, generator=g) * 32768).round().clamp(-32768, 32767).to(torch.int16).cuda()
        _, t_out = eng.out_frames(cfg.chunk_feat_frames)
        y = torch.empty(S, t_out, cfg.llm_dim, device="cuda")
        for i in range(20):
            eng.stream_step(ids, pcm[i % 4], 1.0, adapter_out=y, want_enc=False)
        eng.set_option("use_graph", 0)                      # the profiled step: eager launches, one kernel per ncu result
        eng.stream_step(ids, pcm[0], 1.0, adapter_out=y, want_enc=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eng.stream_step(ids, pcm[1], 1.0, adapter_out=y, want_enc=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
        eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=2)
        g = torch.Generator().manual_seed(7)
        pcm = (0.05 * torch.randn(B, 30 * cfg.sample_rate, generator=g) * 32768).round().clamp(-32768, 32767).to(torch.int16).cuda()

        def one():
            feats = eng.fbank_offline(pcm, 1.0)
            eng.encode_offline(feats, np.full((B,), feats.shape[1], dtype=np.int32), cfg.chunk_size, cfg.left_chunks)
        one()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        one()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    eng.close()
    print("ok")


if __name__ == "__main__":
    main()
