"""Batched session scheduler with the VAD-gated streaming frontend (SURVEY 8 f1 + f3).

The reference runs one feature-gating thread and one predictor thread PER SESSION, hands fbank blocks between them
as Python lists and calls the encoder with batch 1 (bin/dialog_state_pred.py:600-684 feature gating, :719-775 the
processing loop, :793-814 the encoder call; bin/pool.py:61-91 the pipeline pool).  Here one scheduler per GPU gathers
the chunks that arrived during a tick (<= 160 ms), runs ONE batched fbank call for all of them, applies the gating rule
on the device and drains the per-session feature queues with batched encoder+adapter steps:

  gating rule (models/AudioFeatureGating.py:77-109, process_and_gate):
    * features are ALWAYS extracted, so the sample carry / context ring stay continuous;
    * status None (no speech):       the block only enters the session's history ring (last `history_chunks` blocks);
    * status 'ipu_sl' (speech onset): the last `onset_chunks` history blocks are queued in front of the current one
      (dialog_state_pred.py:639-670), each of them is one encoder step;
    * any other status:              the block is queued.
  Each queued block is one `encoder.infer` + adapter step of that session (dialog_state_pred.py:793-814).

Per drain round the head block of every session with a non-empty queue forms one batch; the batch is padded to a
multiple of `bucket` with scratch sessions so that a handful of captured CUDA graphs serve every active-set size.
State (KV rings, adapter cache, sample carry, feature ring, history ring, queued blocks) never leaves the device.
"""
from collections import deque
from typing import Any, Deque, Dict, Hashable, List, NamedTuple, Optional, Tuple

import numpy as np
import torch

IPU_START = "ipu_sl"
IPU_CONT = "ipu_cl"


class Block(NamedTuple):
    """One encoder + adapter step of one session, as `tick` hands it to the LLM stage.  `status` is the label the reference's
    feature-gating thread gives the block (bin/dialog_state_pred.py:639-670): the FIRST block of a speech onset carries
    'ipu_sl' -- which is where AudioLLM.recognize puts the chat prefix (models/audioLLM.py:404-406) -- every other one 'ipu_cl'."""
    enc: torch.Tensor                 # (t, D) encoder frames
    emb: Optional[torch.Tensor]       # (t_out, E) adapter embeddings; with a Handoff: the fp16 LLM input rows, chat prefix included
    status: str                       # when status == 'ipu_sl' (views of the Handoff buffers)
    mask: Optional[torch.Tensor] = None   # with a Handoff: the attention-mask entries of those rows


class Handoff:
    """Pre-allocated LLM-side buffers for `StreamScheduler.tick(handoff=...)` (models/audioLLM.py:383-411): `embeds`
    (capacity, prefix_len + t_out, E) fp16 whose first prefix_len rows of every block hold the chat-prefix embeddings,
    `attn_mask` (capacity, prefix_len + t_out) uint8 and `row_start` (capacity) int32.  One block per encoder step of a tick."""

    def __init__(self, engine, capacity: int, prefix_embeds: torch.Tensor, prefix_mask: Optional[torch.Tensor] = None):
        dev = engine.torch_device
        self.prefix_len = int(prefix_embeds.shape[-2])
        _, t_out = engine.out_frames(engine.cfg.chunk_feat_frames)
        self.rows = self.prefix_len + t_out
        self.embeds = torch.zeros(capacity, self.rows, engine.cfg.llm_dim, dtype=torch.float16, device=dev)
        self.embeds[:, :self.prefix_len] = prefix_embeds.reshape(self.prefix_len, -1).to(dev, torch.float16)
        self.prefix_mask = None if prefix_mask is None else prefix_mask.reshape(-1).to(dev, torch.uint8).contiguous()
        self.attn_mask = torch.zeros(capacity, self.rows, dtype=torch.uint8, device=dev)
        self.row_start = torch.zeros(capacity, dtype=torch.int32, device=dev)


def onset_statuses(n_history: int, status: str) -> List[str]:
    """Labels of the blocks one pushed chunk turns into (bin/dialog_state_pred.py:639-670): the replayed history blocks are
    'ipu_sl', 'ipu_cl', 'ipu_cl', ... and the current block is 'ipu_cl' when history was replayed, else keeps its status."""
    if status != IPU_START or n_history == 0:
        return [status]
    return [IPU_START] + [IPU_CONT] * n_history


def plan_rounds(queue_lens: Dict[Hashable, int], max_batch: int) -> List[List[Hashable]]:
    """Drain order for per-session FIFO queues: every round takes the head block of each session that still has one
    (a session's blocks must run in order, so it appears at most once per round), `max_batch` sessions per batch.
    Pure function of the queue lengths -> testable without a GPU."""
    left = dict(queue_lens)
    rounds: List[List[Hashable]] = []
    while True:
        ready = [k for k, n in left.items() if n > 0]
        if not ready:
            return rounds
        for i in range(0, len(ready), max_batch):
            rounds.append(ready[i:i + max_batch])
        for k in ready:
            left[k] -= 1


def bucket_size(n: int, bucket: int) -> int:
    return (n + bucket - 1) // bucket * bucket


class StreamScheduler:
    def __init__(self, engine, history_chunks: int = 10, onset_chunks: int = 6, bucket: int = 16,
                 max_batch: Optional[int] = None, max_sessions: Optional[int] = None):
        assert 0 <= onset_chunks <= history_chunks
        self.eng, self.cfg = engine, engine.cfg
        self.history_chunks, self.onset_chunks, self.bucket = history_chunks, onset_chunks, max(1, bucket)
        self.max_batch = max_batch or 1 << 30
        self.keys: Dict[Hashable, int] = {}                    # session key -> slot
        self.rows: Dict[Hashable, int] = {}                    # session key -> row of the history tensor
        self.free_rows: List[int] = []
        self.pending: Dict[Hashable, Tuple[Any, Optional[str]]] = {}
        self.queues: Dict[Hashable, Deque[Tuple[torch.Tensor, str]]] = {}
        # the scratch sessions that pad a batch to its bucket come out of the engine's slot pool: the scheduler can only
        # promise what is left of it
        free = engine.max_sessions - engine.stats()["sessions_in_use"] - (self.bucket - 1)
        if free < 1:
            raise ValueError("engine has %d session slots, %d in use: no room for bucket=%d (needs bucket-1 scratch sessions + 1)"
                             % (engine.max_sessions, engine.stats()["sessions_in_use"], self.bucket))
        cap = free if max_sessions is None else int(max_sessions)
        if cap > free:
            raise ValueError("max_sessions=%d exceeds the engine's free slots (%d of %d) minus %d scratch sessions"
                             % (cap, free + self.bucket - 1, engine.max_sessions, self.bucket - 1))
        self.scratch = engine.alloc(self.bucket - 1) if self.bucket > 1 else np.zeros(0, np.int32)
        self._hist = torch.zeros(cap, history_chunks, self.cfg.chunk_feat_frames, self.cfg.feat_dim, device=engine.torch_device)
        self.free_rows = list(range(cap - 1, -1, -1))
        self.stats = {"ticks": 0, "fbank_calls": 0, "encode_calls": 0, "session_steps": 0, "padded_steps": 0}
        # tests: when a list, every block the gating rule queues is appended as (key, feature block, status)
        self.block_log: Optional[List[Tuple[Hashable, torch.Tensor, str]]] = None

    # ---- sessions ---------------------------------------------------------------------------------
    def open(self, key: Hashable) -> int:
        """New session == encoder_cache None, adapter_cache None, pe_index 0 (dialog_state_pred.py:221-232) and a
        zeroed frontend (AudioFeatureGating.reset)."""
        assert key not in self.keys, "session already open"
        if not self.free_rows:
            raise RuntimeError("scheduler is full (max_sessions=%d)" % self._hist.shape[0])
        slot = int(self.eng.alloc(1)[0])
        self.keys[key], self.rows[key] = slot, self.free_rows.pop()
        self._hist[self.rows[key]].zero_()
        self.queues[key] = deque()
        return slot

    def close(self, key: Hashable) -> None:
        self.eng.free([self.keys.pop(key)])
        self.free_rows.append(self.rows.pop(key))
        self.queues.pop(key)
        self.pending.pop(key, None)

    def reset(self, key: Hashable) -> None:
        self.eng.reset([self.keys[key]])
        self._hist[self.rows[key]].zero_()
        self.queues[key].clear()
        self.pending.pop(key, None)

    # ---- data path ----------------------------------------------------------------------------------
    def push(self, key: Hashable, pcm, status: Optional[str]) -> None:
        """One chunk of `samples_per_chunk` samples (int16 or float32, host or device) with its VAD annotation."""
        assert key in self.keys and key not in self.pending, "one chunk per session per tick"
        self.pending[key] = (pcm, status)

    def tick(self, scale: Optional[float] = None, handoff: Optional[Handoff] = None) -> Dict[Hashable, List[Block]]:
        """Process everything pushed since the last tick.  Returns, per session that produced output, the list of
        Block(encoder frames (t, D), adapter embeddings (t_out, E), status) in stream order (one entry per queued block).
        `scale` None = the engine's rule: 1.0 for int16 PCM, cfg.pcm_scale for float audio in [-1, 1]."""
        self.stats["ticks"] += 1
        out: Dict[Hashable, List[Block]] = {}
        if self.pending:
            keys = list(self.pending)
            dev = self.eng.torch_device
            pcm = torch.stack([torch.as_tensor(self.pending[k][0]).to(dev, non_blocking=True) for k in keys])
            ids = np.array([self.keys[k] for k in keys], np.int32)
            feats = self.eng.fbank_stream(ids, pcm, scale)                       # carry + context ring advance for everyone
            self.stats["fbank_calls"] += 1
            silent = [i for i, k in enumerate(keys) if self.pending[k][1] is None]
            if silent:                                                           # history ring: drop the oldest, append
                rows = torch.tensor([self.rows[keys[i]] for i in silent], device=dev)
                h = self._hist.index_select(0, rows)
                h = torch.cat([h[:, 1:], feats.index_select(0, torch.tensor(silent, device=dev)).unsqueeze(1)], dim=1)
                self._hist.index_copy_(0, rows, h)
            for i, k in enumerate(keys):
                status = self.pending[k][1]
                if status is None:
                    continue
                blocks = [feats[i]]
                if status == IPU_START and self.onset_chunks > 0:
                    hist = self._hist[self.rows[k], self.history_chunks - self.onset_chunks:].clone()
                    blocks = list(hist.unbind(0)) + blocks
                labelled = list(zip(blocks, onset_statuses(len(blocks) - 1, status)))
                self.queues[k].extend(labelled)
                if self.block_log is not None:
                    self.block_log.extend((k, b, lab) for b, lab in labelled)
            self.pending.clear()
        cursor = 0
        for batch in plan_rounds({k: len(q) for k, q in self.queues.items()}, self.max_batch):
            heads = [self.queues[k].popleft() for k in batch]
            blocks = [h[0] for h in heads]
            n = len(batch)
            pad = bucket_size(n, self.bucket) - n
            ids = np.concatenate([np.array([self.keys[k] for k in batch], np.int32), self.scratch[:pad]]).astype(np.int32)
            x = torch.stack(blocks + [blocks[0]] * pad)
            if handoff is not None:
                # fused hand-off: rows of this batch land in the next free blocks of the caller's buffers (audioLLM.py:404-411)
                n_all = n + pad
                if cursor + n_all > handoff.embeds.shape[0]:
                    raise RuntimeError("Handoff capacity %d exceeded in one tick" % handoff.embeds.shape[0])
                onset = [h[1] == IPU_START for h in heads] + [False] * pad
                self.eng.handoff_arm(n_all, handoff.embeds[cursor:], handoff.prefix_len, onset, handoff.prefix_mask,
                                     handoff.attn_mask[cursor:], handoff.row_start[cursor:])
                enc, _ = self.eng.encode_stream(ids, x, want_adapter=False)
                self.stats["encode_calls"] += 1
                self.stats["session_steps"] += n
                self.stats["padded_steps"] += pad
                for i, k in enumerate(batch):
                    st0 = 0 if heads[i][1] == IPU_START else handoff.prefix_len
                    out.setdefault(k, []).append(Block(enc[i], handoff.embeds[cursor + i, st0:], heads[i][1],
                                                       handoff.attn_mask[cursor + i, st0:]))
                cursor += n_all
                continue
            enc, emb = self.eng.encode_stream(ids, x)
            self.stats["encode_calls"] += 1
            self.stats["session_steps"] += n
            self.stats["padded_steps"] += pad
            for i, k in enumerate(batch):
                out.setdefault(k, []).append(Block(enc[i], emb[i] if emb is not None else None, heads[i][1]))
        return out

    def history(self, key: Hashable) -> torch.Tensor:
        return self._hist[self.rows[key]]
