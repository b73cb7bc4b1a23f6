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
        self.max_sessions = 64

    def stats(self):
        return {"sessions_in_use": self.next_slot}

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
    # labels of the replayed blocks (bin/dialog_state_pred.py:639-670): the first carries the onset, the rest continue
    assert [e.status for e in out["a"]] == ["ipu_sl", "ipu_cl", "ipu_cl"]
    assert len(out["b"]) == 1 and out["b"][0].status == "ipu_cl"
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


def test_onset_without_history_keeps_its_label_and_capacity_is_checked():
    from freeze_omni_b200.scheduler import onset_statuses
    assert onset_statuses(0, "ipu_sl") == ["ipu_sl"] and onset_statuses(3, "ipu_cl") == ["ipu_cl"]
    assert onset_statuses(6, "ipu_sl") == ["ipu_sl"] + ["ipu_cl"] * 6
    eng = FakeEngine()
    sch = StreamScheduler(eng, history_chunks=4, onset_chunks=0, bucket=4)
    sch.open("a")
    sch.push("a", np.zeros(2560, np.int16), "ipu_sl")
    assert [b.status for b in sch.tick()["a"]] == ["ipu_sl"]
    # default capacity = the engine's free slots minus the scratch sessions (open() used to fail at the 50th session)
    eng2 = FakeEngine()
    eng2.max_sessions = 8
    sch2 = StreamScheduler(eng2, bucket=4)
    assert sch2._hist.shape[0] == 8 - 3
    import pytest
    with pytest.raises(ValueError):
        StreamScheduler(FakeEngine(), bucket=4, max_sessions=64)


class OracleFrontEngine(FakeEngine):
    """FakeEngine whose frontend is the CPU oracle's stateful fbank, one per slot: the scheduler's gating logic runs on the
    same feature blocks the reference's AudioFeatureGating produced for the golden."""

    def __init__(self):
        super().__init__()
        from oracle import freeze_omni_oracle as O
        self.O, self.front = O, {}

    def fbank_stream(self, ids, pcm, scale=None):
        self.fbank_calls.append(list(ids))
        out = []
        for i, s in enumerate(ids):
            fr = self.front.setdefault(int(s), self.O.StreamingFrontend())
            out.append(fr.process(torch.as_tensor(pcm[i]).float(), 1.0 if scale is None else scale))
        return torch.cat(out, 0)


def test_gating_rule_pinned_to_reference_process_and_gate():
    """tests/golden/gating.npz holds what models/AudioFeatureGating.process_and_gate + the relabelling loop of
    bin/dialog_state_pred.py:626-670 queued for a seeded stream (generated by tests/golden/make_golden.py from the reference
    module itself): same blocks, same order, same labels, same history ring."""
    from conftest import load_golden
    g = load_golden("gating")
    eng = OracleFrontEngine()
    sch = StreamScheduler(eng, history_chunks=10, onset_chunks=6, bucket=1, max_sessions=4)
    sch.open("u")
    sch.block_log = []
    per_tick = []
    for i, st in enumerate(g["statuses"]):
        a = g["pcm"][i * 2560:(i + 1) * 2560].astype(np.float32) / 32768.0
        sch.push("u", a, str(st) if str(st) else None)
        out = sch.tick(float(g["scale"]))
        per_tick.append(len(out.get("u", [])))
    assert len(sch.block_log) == len(g["labels"])
    assert [lab for _, _, lab in sch.block_log] == [str(x) for x in g["labels"]]
    got = torch.stack([b for _, b, _ in sch.block_log]).numpy()
    err = np.abs(got - g["blocks"]) / np.maximum(np.abs(g["blocks"]), 1.0)
    assert err.max() < 1e-5, err.max()
    herr = np.abs(sch.history("u").numpy() - g["history"]) / np.maximum(np.abs(g["history"]), 1.0)
    assert herr.max() < 1e-5
    # blocks per tick: the reference emits 1 + 6 at an onset, 1 while speaking, 0 when silent
    want = [int((g["owner"] == i).sum()) for i in range(len(g["statuses"]))]
    assert per_tick == want
