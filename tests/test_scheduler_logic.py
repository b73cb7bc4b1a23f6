"""Host logic of the batched session scheduler (no GPU): drain order of the per-session queues and bucket padding,
plus the gating rule driven through a fake engine that records the calls."""
import numpy as np
import torch

from freeze_omni_b200.scheduler import StreamScheduler, bucket_size, plan_rounds


def test_plan_rounds_keeps_session_order_and_batches():
    rounds = plan_rounds({"a": 1, "b": 7, "c": 0, "d": 2}, max_batch=8)
    assert rounds[0] == ["a", "b", "d"] and rounds[1] == ["b", "d"] and rounds[2:] == [["b"]] * 5
    # a session never appears twice in one batch; max_batch splits a round
    rounds = plan_rounds({i: 1 for i in range(5)}, max_batch=2)
    assert rounds == [[0, 1], [2, 3], [4]]
    assert plan_rounds({}, 4) == [] and plan_rounds({"x": 0}, 4) == []


def test_bucket_size():
    assert [bucket_size(n, 16) for n in (1, 16, 17, 33)] == [16, 16, 32, 48]
    assert bucket_size(5, 1) == 5


class FakeCfg:
    chunk_feat_frames, feat_dim, samples_per_chunk = 19, 80, 2560


class FakeEngine:
    """Records what the scheduler asks for; `fbank_stream` returns blocks tagged with (slot, call index)."""

    def __init__(self):
        self.cfg, self.torch_device = FakeCfg(), torch.device("cpu")
        self.next_slot, self.fbank_calls, self.encode_calls = 0, [], []

    def alloc(self, n):
        ids = np.arange(self.next_slot, self.next_slot + n, dtype=np.int32)
        self.next_slot += n
        return ids

    def free(self, ids):
        pass

    def reset(self, ids):
        pass

    def fbank_stream(self, ids, pcm, scale=None):
        self.fbank_calls.append(list(ids))
        tag = torch.tensor([100.0 * int(s) + len(self.fbank_calls) for s in ids], dtype=torch.float32)
        return tag.view(-1, 1, 1).expand(len(ids), 19, 80).clone()

    def encode_stream(self, ids, feats):
        self.encode_calls.append((list(ids), feats[:, 0, 0].tolist()))
        return feats[:, :4, :8].clone(), feats[:, :2, :4].clone()


def test_gating_rule_and_queue_order():
    eng = FakeEngine()
    sch = StreamScheduler(eng, history_chunks=4, onset_chunks=2, bucket=4, max_sessions=8)
    a, b = sch.open("a"), sch.open("b")
    pcm = np.zeros(2560, np.int16)
    # ticks 1-3: a silent (history fills), b speaking
    for _ in range(3):
        sch.push("a", pcm, None)
        sch.push("b", pcm, "ipu_cl")
        out = sch.tick()
        assert list(out) == ["b"] and len(out["b"]) == 1
    assert sch.history("a")[:, 0, 0].tolist() == [0.0, 100.0 * a + 1, 100.0 * a + 2, 100.0 * a + 3]
    # tick 4: onset for a -> last 2 history blocks, then the current one; b keeps going (one block)
    sch.push("a", pcm, "ipu_sl")
    sch.push("b", pcm, "ipu_cl")
    out = sch.tick()
    assert [float(e[0][0, 0]) for e in out["a"]] == [100.0 * a + 2, 100.0 * a + 3, 100.0 * a + 4]
    assert len(out["b"]) == 1
    # every fbank call covered both sessions (features are always extracted); batches were padded to the bucket
    assert all(c == [a, b] for c in eng.fbank_calls)
    assert all(len(ids) % 4 == 0 for ids, _ in eng.encode_calls)
    first = eng.encode_calls[-3]
    assert first[0][:2] == [a, b] and set(first[0][2:]) <= set(sch.scratch.tolist())
    assert eng.encode_calls[-2][0][0] == a and eng.encode_calls[-1][0][0] == a
    # the history ring is not touched while speaking
    assert sch.history("a")[-1, 0, 0].item() == 100.0 * a + 3
    assert sch.stats["session_steps"] == 3 + 4 and sch.stats["ticks"] == 4
    sch.close("a")
    sch.close("b")
