#!/usr/bin/env python
"""In-chain timeline of ONE graph-replayed streaming step, from per-CTA %globaltimer records (development build:
FO_TRACE_BUILD=1 python -m freeze_omni_b200.build  ->  freeze_omni_b200/libfo_b200_trace.so).

Unlike an ncu launch list (serialised, cold) this shows the step as it runs: when the CTAs of each kernel start (PDL lets
them start while the predecessor drains), when their dependency wait returns, when the first operands have landed, when the
accumulator is complete and when the CTA exits.

    FO_B200_LIB=freeze_omni_b200/libfo_b200_trace.so python tools/step_timeline.py [--sessions 64] [--opt name=value]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FO_B200_LIB", os.path.join(ROOT, "freeze_omni_b200", "libfo_b200_trace.so"))
from freeze_omni_b200 import _lib  # noqa: E402
from freeze_omni_b200.config import load_path_config  # noqa: E402
from freeze_omni_b200.engine import Engine  # noqa: E402
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state  # noqa: E402

KNAME = {1: "gemm_tc", 2: "attention_stream", 3: "layer_norm_reduce"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=64)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--plan", action="append", default=[], help="N:K:swap:bn:split[:cap_kb]")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    cfg = load_path_config("shipped")
    S = args.sessions
    eng = Engine(cfg, make_encoder_state(cfg, 0), make_adapter_state(cfg, 0), dtype=torch.bfloat16, max_sessions=S,
                 max_stream_frames=cfg.chunk_feat_frames)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    for pl in args.plan:
        f = [int(x) for x in pl.split(":")] + [0]
        eng.set_option("tc_plan", f[0] | (f[1] << 16) | (f[2] << 32) | (f[3] << 33) | (f[4] << 42) | (f[5] << 48))
    ids = eng.alloc(S)
    g = torch.Generator().manual_seed(5)
    pcm = (0.05 * torch.randn(4, S, cfg.samples_per_chunk, generator=g) * 32768).round().clamp(-32768, 32767).to(torch.int16).cuda()
    _, t_out = eng.out_frames(cfg.chunk_feat_frames)
    y = torch.empty(S, t_out, cfg.llm_dim, device="cuda")
    for i in range(24):                                        # windows full, graph captured and replayed
        eng.stream_step(ids, pcm[i % 4], 1.0, adapter_out=y, want_enc=False)
    torch.cuda.synchronize()
    eng.set_option("trace", 60000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.stream_step(ids, pcm[0], 1.0, adapter_out=y, want_enc=False)
    e1.record()
    torch.cuda.synchronize()
    buf = np.zeros((60000, 8), dtype=np.uint64)
    n = C.c_int64()
    _lib.check(eng.lib.fo_debug_trace_read(eng._h, buf.ctypes.data, 60000, C.byref(n)))
    eng.set_option("trace", 0)
    rec = buf[:n.value]
    if len(rec) == 0:
        raise SystemExit("no trace records: is FO_B200_LIB a FO_TRACE_BUILD library?")
    # kid 9: k-block arrival times at the MMA thread of one CTA per GEMM launch (three records of six stamps each)
    kb = rec[(rec[:, 0] & 0xFF) == 9]
    iss = rec[(rec[:, 0] & 0xFF) == 10]
    rec = rec[((rec[:, 0] & 0xFF) != 9) & ((rec[:, 0] & 0xFF) != 10)]
    if len(iss):
        v = iss[:, 2:8].astype(np.int64)
        ok = (v > 0).all(axis=1)
        v = v[ok]
        print("MMA thread, k-blocks 4 and 5 of one CTA per launch (ns, median over %d launches): wait-done -> 4 MMAs issued %d / %d, "
              "-> commit issued %d / %d, commit -> next wait-done %d" % (len(v), np.median(v[:, 1] - v[:, 0]), np.median(v[:, 4] - v[:, 3]),
              np.median(v[:, 2] - v[:, 1]), np.median(v[:, 5] - v[:, 4]), np.median(v[:, 3] - v[:, 2])))
    kb_by = {}
    for r in kb:
        part = int((r[0] >> 16) & 0xFFFF)
        grid = (int(r[1] & 0xFFFFF), int((r[1] >> 20) & 0xFFFFF), int(r[1] >> 40))
        kb_by.setdefault(grid, {}).setdefault(part, []).append(r[2:8].astype(np.int64))
    print("k-block arrival deltas (ns) at the MMA thread of one CTA, median over the launches of each grid:")
    for grid, parts in kb_by.items():
        nl = min(len(v) for v in parts.values())
        if nl == 0 or len(parts) < 3:
            continue
        seqs = np.stack([np.concatenate([parts[p_][i] for p_ in (0, 1, 2)]) for i in range(nl)])
        valid = (seqs > 0).all(axis=0)
        d = np.diff(seqs[:, valid], axis=1)
        print("  grid %-12s launches %3d: %s" % ("x".join(map(str, grid)), nl, " ".join("%d" % x for x in np.median(d, axis=0))))
    kid = (rec[:, 0] & 0xFF).astype(np.int64)
    smid = ((rec[:, 0] >> 8) & 0xFF).astype(np.int64)
    aux = ((rec[:, 0] >> 16) & 0xFFFF).astype(np.int64)
    gx = (rec[:, 1] & 0xFFFFF).astype(np.int64)
    gy = ((rec[:, 1] >> 20) & 0xFFFFF).astype(np.int64)
    gz = (rec[:, 1] >> 40).astype(np.int64)
    t = rec[:, 2:8].astype(np.int64)
    t0 = t[:, 0].min()
    t = t - t0
    order = np.argsort(t[:, 0], kind="stable")
    # launches: consecutive (in start order) records with the same kernel id + grid
    launches, cur = [], None
    for i in order:
        key = (kid[i], gx[i], gy[i], gz[i], aux[i])
        if cur is None or cur["key"] != key or len(cur["idx"]) >= gx[i] * gy[i] * gz[i]:
            cur = {"key": key, "idx": []}
            launches.append(cur)
        cur["idx"].append(i)
    rows = []
    prev_end = 0
    for L in launches:
        ix = np.asarray(L["idx"])
        k, x, yy, z, a = L["key"]
        st, wt, f1, acc, wr, en = (t[ix, j] for j in range(6))
        rows.append({"kernel": KNAME.get(int(k), str(k)), "grid": [int(x), int(yy), int(z)], "n_out": int(a) * 64 if k == 1 else None,
                     "ctas": int(len(ix)), "sms": int(len(set(smid[ix].tolist()))),
                     "first_start_us": st.min() / 1e3, "last_start_us": st.max() / 1e3,
                     "dep_wait_done_us_med": float(np.median(wt)) / 1e3 if (wt > 0).any() else None,
                     "first_operands_us_med": float(np.median(f1)) / 1e3, "acc_done_us_med": float(np.median(acc)) / 1e3,
                     "last_end_us": en.max() / 1e3, "cta_span_us_med": float(np.median(en - st)) / 1e3,
                     "since_prev_end_us": (en.max() - prev_end) / 1e3})
        prev_end = max(prev_end, en.max())
    step_us = e0.elapsed_time(e1) * 1e3
    print("step %.1f us (CUDA events, with trace stamps); %d CTA records, %d launches traced" % (step_us, len(rec), len(rows)))
    print("%-18s %-14s %5s %4s | %8s %8s %8s %8s %8s %8s | %7s %7s" % ("kernel", "grid", "ctas", "sms", "start0", "startN", "depwait", "operand", "accdone",
                                                                  "end", "ctaspan", "+prev"))
    for r in rows:
        print("%-18s %-14s %5d %4d | %8.2f %8.2f %8s %8.2f %8.2f %8.2f | %7.2f %7.2f" % (
            r["kernel"] + ("/%d" % r["n_out"] if r["n_out"] else ""), "x".join(map(str, r["grid"])), r["ctas"], r["sms"], r["first_start_us"],
            r["last_start_us"], "%.2f" % r["dep_wait_done_us_med"] if r["dep_wait_done_us_med"] is not None else "-",
            r["first_operands_us_med"], r["acc_done_us_med"], r["last_end_us"], r["cta_span_us_med"], r["since_prev_end_us"]))
    # per-kernel-class averages of the increment each launch adds to the chain
    agg = {}
    for r in rows:
        key = (r["kernel"], tuple(r["grid"]), r["n_out"])
        agg.setdefault(key, []).append(r)
    print("\nper kernel class: launches, mean chain increment (end - previous end), mean CTA span, mean (dep-wait-done - first start)")
    for key, lst in sorted(agg.items(), key=lambda kv: -sum(r["since_prev_end_us"] for r in kv[1])):
        inc = np.mean([r["since_prev_end_us"] for r in lst])
        span = np.mean([r["cta_span_us_med"] for r in lst])
        pre = np.mean([(r["dep_wait_done_us_med"] or r["first_start_us"]) - r["first_start_us"] for r in lst])
        print("  %-18s %-12s n_out=%-5s n=%3d  increment %6.2f us  cta span %6.2f us  prologue before dep %5.2f us  total %7.1f us" % (
            key[0], "x".join(map(str, key[1])), key[2], len(lst), inc, span, pre, inc * len(lst)))
    if args.out:
        json.dump({"step_us": step_us, "launches": rows}, open(args.out, "w"))
    eng.close()


if __name__ == "__main__":
    main()
