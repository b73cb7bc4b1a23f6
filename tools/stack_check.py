"""Weight-streaming layer stack (csrc/fo_stack.cu) against the per-kernel chain on the shipped bf16 context: same PCM
through two groups of sessions, one per execution form (option stack_rows), max-abs difference of the encoder / adapter
outputs per step, then the device time per step of both forms (graph replay, CUDA events).
    python tools/stack_check.py [sessions ...]          FO_STACK_TRACE=1 prints the per-phase stamps of the eager launches
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402


def main():
    ns = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
    cfg = load_path_config(os.environ.get("FO_CFG", "shipped"))
    eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=16)
    g = torch.Generator().manual_seed(5)
    st = torch.cuda.current_stream()
    for n in ns:
        ia, ib = eng.alloc(n), eng.alloc(n)
        worst_e = worst_y = 0.0
        for i in range(24):
            pcm = (0.05 * torch.randn(n, cfg.samples_per_chunk, generator=g) * 32768).round().clamp(-32768, 32767).to(torch.int16).cuda()
            eng.set_option("stack_rows", 16)
            e1, y1 = eng.stream_step(ia, pcm, 1.0)
            eng.set_option("stack_rows", 0)
            e0, y0 = eng.stream_step(ib, pcm, 1.0)
            worst_e = max(worst_e, float((e1 - e0).abs().max()))
            worst_y = max(worst_y, float((y1 - y0).abs().max()))
            assert torch.isfinite(e1).all() and torch.isfinite(y1).all(), "non-finite output at step %d" % i
        times = {}
        for rows in (16, 0):
            eng.set_option("stack_rows", rows)
            ids = ia if rows else ib
            for _ in range(10):
                eng.stream_step(ids, pcm, 1.0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(200):
                eng.stream_step(ids, pcm, 1.0, want_enc=False)
            b.record(st)
            torch.cuda.synchronize()
            times[rows] = a.elapsed_time(b) / 200
        print("sessions %d: stack vs chain max-abs encoder %.3g adapter %.3g; ms/step stack %.4f chain %.4f; stack launches %d; saturations %d"
              % (n, worst_e, worst_y, times[16], times[0], eng.get_option("stack_launches"), eng.stats().get("act_saturations", -1)), flush=True)
        eng.free(ia)
        eng.free(ib)
    eng.set_option("stack_rows", 16)
    eng.close()


if __name__ == "__main__":
    main()
