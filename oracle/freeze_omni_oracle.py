"""CPU oracle for the streaming speech encoder + adapter path  --  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a state dict, what the reference's PyTorch modules
compute on the hot path (fbank frontend -> CMVN -> Conv2d x4 subsampling -> rel-pos chunk
transformer with KV cache -> CNN adapter with conv cache).  It is the checker for the CUDA
path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``freeze_omni_b200/`` imports it and the
product path has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own, so
the pin is made by executing the reference modules themselves (imported from /root/reference
in the authoring container by ``tests/golden/make_golden.py``) on the seeded state dict of
``freeze_omni_b200/weights.py`` and committing their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.  The fbank
restates the third-party ``torchaudio.compliance.kaldi.fbank`` (pinned torchaudio==2.2.0 in the
reference's requirements.txt:9; 2.11.0 in this image, kaldi.py:514-645) and is pinned the same
way plus the known-answer values recorded in SURVEY 4.

Arithmetic is fp32 on torch CPU ops, like the reference's own CPU path.  Every function cites the
reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

FLT_EPSILON = 1.1920928955078125e-07
MIN_VALUE = -65504.0     # numpy.finfo(float16).min, attention.py:288


# ------------------------------------------------------------------------------------------------
# Frontend: torchaudio.compliance.kaldi.fbank as called at bin/inference.py:77-78 and
# models/AudioFeatureGating.py:65-69 (dither=0, povey window, snip_edges, power, log, no energy)
# ------------------------------------------------------------------------------------------------
def povey_window(window_size: int) -> Tensor:
    """kaldi.py:98-100: hann_window(periodic=False) ** 0.85."""
    return torch.hann_window(window_size, periodic=False, dtype=torch.float32).pow(0.85)


def mel_banks(num_bins: int, padded_window: int, sample_freq: float, low_freq: float = 20.0,
              high_freq: float = 0.0) -> Tensor:
    """kaldi.py:436-511 without VTLN.  Returns (num_bins, padded_window//2 + 1); the Nyquist
    column is the zero pad added at kaldi.py:627."""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    n_fft_bins = padded_window // 2
    bin_width = sample_freq / padded_window
    mel_lo = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_hi - mel_lo) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (bin_width * torch.arange(float(n_fft_bins))) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return F.pad(bins, (0, 1), value=0.0).float()


def fbank(wave: Tensor, sample_rate: int = 16000, frame_length_ms: float = 25.0,
          frame_shift_ms: float = 10.0, num_mel_bins: int = 80, preemph: float = 0.97) -> Tensor:
    """wave: (N,) fp32 in int16 range.  Returns (m, num_mel_bins), m = 1 + (N - win)//shift
    (kaldi.py:64-83 snip_edges).  Steps: kaldi.py:183-186 (DC), :193-199 (pre-emphasis with
    replicate pad), :202-205 (window), :207-212 (zero pad to 2^k), :616-618 (|rfft|^2),
    :630-633 (mel matmul, log floor)."""
    win = int(sample_rate * frame_length_ms * 0.001)
    shift = int(sample_rate * frame_shift_ms * 0.001)
    padded = 1 << (win - 1).bit_length()
    n = wave.numel()
    if n < win:
        return torch.empty(0, num_mel_bins)
    m = 1 + (n - win) // shift
    frames = wave.float().as_strided((m, win), (shift, 1))
    frames = frames - frames.mean(dim=1, keepdim=True)
    prev = torch.cat([frames[:, :1], frames[:, :-1]], dim=1)
    frames = frames - preemph * prev
    frames = frames * povey_window(win).unsqueeze(0)
    frames = F.pad(frames, (0, padded - win))
    power = torch.fft.rfft(frames).abs().pow(2.0)
    mel = torch.mm(power, mel_banks(num_mel_bins, padded, float(sample_rate)).T)
    return torch.max(mel, torch.tensor(FLT_EPSILON)).log()


class StreamingFrontend:
    """Sample carry + feature-context ring of bin/inference.py:43-80 (audioEncoderProcessor) ==
    models/AudioFeatureGating.py:43-75 with its own constants.  ``process`` takes one chunk of
    ``samples_per_chunk`` float samples (already scaled by the caller's choice of ``scale``,
    SURVEY 2.4-9) and returns the (1, context + frames_per_chunk, feat) block fed to the encoder."""

    def __init__(self, sample_rate=16000, frame_length_ms=25, frame_shift_ms=10, frames_per_chunk=16,
                 context_frames=3, feat_dim=80):
        self.sr, self.fl, self.fs = sample_rate, frame_length_ms, frame_shift_ms
        self.win = sample_rate * frame_length_ms // 1000
        self.shift = sample_rate * frame_shift_ms // 1000
        self.carry = self.win - self.shift
        self.chunk = self.shift * frames_per_chunk
        self.n_new, self.n_ctx, self.feat = frames_per_chunk, context_frames, feat_dim
        self.reset()

    def reset(self):
        self.samples = torch.zeros(self.carry + self.chunk)
        self.feats = torch.zeros(self.n_ctx + self.n_new, self.feat)

    def process(self, audio: Tensor, scale: float = 32768.0) -> Tensor:
        x = audio.reshape(-1).float() * scale
        assert x.numel() == self.chunk
        self.samples = torch.cat([self.samples[-self.carry:], x])          # inference.py:61-64
        new = fbank(self.samples, self.sr, self.fl, self.fs, self.feat)    # inference.py:77-78
        assert new.shape[0] == self.n_new
        self.feats = torch.cat([self.feats[-self.n_ctx:], new], dim=0)     # inference.py:66-69
        return self.feats.unsqueeze(0).clone()


# ------------------------------------------------------------------------------------------------
# Masks (models/masks.py) -- integer/bool work, bit-exact
# ------------------------------------------------------------------------------------------------
def chunk_window(i: int, size: int, chunk: int, left: int) -> Tuple[int, int]:
    """Key range [start, end) visible to row i: masks.py:50-56."""
    start = 0 if left < 0 else max((i // chunk - left) * chunk, 0)
    end = min((i // chunk + 1) * chunk, size)
    return start, end


def subsequent_chunk_mask(size: int, chunk: int, left: int = -1) -> Tensor:
    """masks.py:23-57, closed form instead of the per-row loop."""
    i = torch.arange(size).unsqueeze(1)
    j = torch.arange(size).unsqueeze(0)
    blk = torch.div(i, chunk, rounding_mode="floor")
    start = torch.zeros_like(i) if left < 0 else torch.clamp((blk - left) * chunk, min=0)
    end = torch.clamp((blk + 1) * chunk, max=size)
    return (j >= start) & (j < end)


def pad_mask(lengths: Tensor, max_len: int) -> Tensor:
    """True where padded; the 2-argument form encoder.py:142 calls (masks.py:125-151 takes one)."""
    return torch.arange(max_len).unsqueeze(0) >= lengths.reshape(-1, 1)


def offline_attention_mask(valid: Tensor, chunk: int, left: int) -> Tensor:
    """masks.py:59-123 (static branch :110-120 and the chunk<=0 fallthrough :121-122).
    valid: (B, 1, T) bool.  Returns (B, T, T) or (B, 1, T)."""
    if chunk > 0:
        return valid & subsequent_chunk_mask(valid.size(2), chunk, left).unsqueeze(0)
    return valid


# ------------------------------------------------------------------------------------------------
# Encoder pieces
# ------------------------------------------------------------------------------------------------
def cmvn(sd: State, x: Tensor) -> Tensor:
    """encoder/cmvn.py:32-34."""
    if "global_cmvn.mean" not in sd:
        return x
    return (x - sd["global_cmvn.mean"]) * sd["global_cmvn.istd"]


def conv_subsample4(sd: State, x: Tensor) -> Tensor:
    """encoder/subsampling.py:58-63: two 3x3 stride-2 convs + ReLU, (b,c,t,f)->(b,t,c*f), Linear."""
    y = F.relu(F.conv2d(x.unsqueeze(1), sd["enc.0.core.conv.0.weight"], sd["enc.0.core.conv.0.bias"], stride=2))
    y = F.relu(F.conv2d(y, sd["enc.0.core.conv.2.weight"], sd["enc.0.core.conv.2.bias"], stride=2))
    b, c, t, f = y.shape
    y = y.permute(0, 2, 1, 3).reshape(b, t, c * f)
    return F.linear(y, sd["enc.0.core.out.0.weight"], sd["enc.0.core.out.0.bias"])


def embed(sd: State, x: Tensor) -> Tensor:
    """encoder/transformer.py:186-200: Linear, LayerNorm, (Dropout), ReLU -- or identity."""
    if "enc.1.embed.0.weight" not in sd:
        return x
    d = x.size(-1)
    y = F.linear(x, sd["enc.1.embed.0.weight"], sd["enc.1.embed.0.bias"])
    return F.relu(F.layer_norm(y, (d,), sd["enc.1.embed.1.weight"], sd["enc.1.embed.1.bias"]))


def div_term(d_model: int) -> Tensor:
    """encoder/attention.py:85-87."""
    return torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))


def pos_table(start: int, length: int, d_model: int) -> Tensor:
    """Rows start..start+length of the sinusoid table: attention.py:111-117 (== :27-34)."""
    pe = torch.zeros(length, d_model)
    pos = torch.arange(start, start + length, dtype=torch.float32).unsqueeze(1)
    dt = div_term(d_model)
    pe[:, 0::2] = torch.sin(pos * dt)
    pe[:, 1::2] = torch.cos(pos * dt)
    return pe


def _split_heads(x: Tensor, h: int) -> Tensor:
    b, t, d = x.shape
    return x.view(b, t, h, d // h).transpose(1, 2)


def attention_scores(sd: State, pfx: str, q: Tensor, k: Tensor, pos_emb: Tensor, h: int) -> Tensor:
    """attention.py:430-450 / :370-390: ((q+u) K^T + (q+v) P^T) / sqrt(d_k), P indexed by key
    position (no rel_shift)."""
    dk = q.size(-1)
    p = _split_heads(F.linear(pos_emb, sd[pfx + "linear_pos.weight"]), h)            # (1,h,n,dk)
    qu = q + sd[pfx + "pos_bias_u"].unsqueeze(1)                                     # (b,h,t,dk)
    qv = q + sd[pfx + "pos_bias_v"].unsqueeze(1)
    return (torch.matmul(qu, k.transpose(-2, -1)) + torch.matmul(qv, p.transpose(-2, -1))) / math.sqrt(dk)


def feed_forward(sd: State, p: str, y: Tensor, cache: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Positionwise layer of one block.  PositionwiseFeedForward (attention.py:137-143): w_2(relu(w_1 y)); or
    Conv1dLinear (attention.py:241-266): causal depthwise Conv1d over time (left context = `cache`, the last k-1
    input frames, zeros at the start -- the left_padding of :251) -> 1x1 Conv1d -> ReLU -> Linear.  The reference's
    streaming wiring of Conv1dLinear is broken, so the cache carry here is the one that makes chunked
    evaluation equal to `forward` on the concatenated input.  Returns (out, new_cache (B, C, k-1) or None)."""
    if p + "feed_forward.w_1.0.weight" in sd:
        wd, bd = sd[p + "feed_forward.w_1.0.weight"], sd[p + "feed_forward.w_1.0.bias"]
        k = wd.size(-1)
        x = y.transpose(1, 2)                                                         # (B, C, T)
        left = cache if cache is not None else x.new_zeros(x.size(0), x.size(1), k - 1)
        x = torch.cat([left, x], dim=2)
        new_cache = x[:, :, -(k - 1):].clone()
        x = F.conv1d(x, wd, bd, groups=x.size(1))
        x = F.conv1d(x, sd[p + "feed_forward.w_1.1.weight"], sd[p + "feed_forward.w_1.1.bias"])
        x = F.relu(x).transpose(1, 2)
        return F.linear(x, sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"]), new_cache
    if sd[p + "feed_forward.w_1.weight"].dim() == 3:
        # MultiLayeredConv1d (attention.py:158-196): Conv1d(k, padding (k-1)//2) -> ReLU -> Conv1d(k, padding (k-1)//2) over time;
        # full-utterance forward only (the module has no infer(), attention.py:145-196)
        w1, w2 = sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_2.weight"]
        pad = (w1.size(-1) - 1) // 2
        x = F.relu(F.conv1d(y.transpose(1, 2), w1, sd[p + "feed_forward.w_1.bias"], padding=pad))
        return F.conv1d(x, w2, sd[p + "feed_forward.w_2.bias"], padding=pad).transpose(1, 2), None
    return F.linear(F.relu(F.linear(y, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"])),
                    sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"]), None


def _attn_residual(sd: State, p: str, x: Tensor, y: Tensor, att: Tensor, concat_after: bool) -> Tensor:
    """x + att, or x + concat_linear(cat(y, att)) with y the (possibly normalised) layer input (transformer.py:85-90,108-116)."""
    if concat_after:
        return x + F.linear(torch.cat((y, att), dim=-1), sd[p + "concat_linear.weight"], sd[p + "concat_linear.bias"])
    return x + att


def layer_stream(sd: State, i: int, x: Tensor, pos_emb: Tensor, kv: Optional[List[Tensor]], h: int,
                 window: int, normalize_before: bool = True, concat_after: bool = False) -> Tuple[Tensor, List[Tensor]]:
    """TransformerLayer.infer (transformer.py:103-130) + MultiHeadedAttention.infer
    (attention.py:407-459) + PositionwiseFeedForward.infer (attention.py:141-143)."""
    p = "enc.1.encoders.%d." % i
    a = p + "self_attn."
    d = x.size(-1)
    y = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"]) if normalize_before else x
    q = _split_heads(F.linear(y, sd[a + "linear_q.weight"], sd[a + "linear_q.bias"]), h)
    k = _split_heads(F.linear(y, sd[a + "linear_k.weight"], sd[a + "linear_k.bias"]), h)
    v = _split_heads(F.linear(y, sd[a + "linear_v.weight"], sd[a + "linear_v.bias"]), h)
    if kv is not None:
        k = torch.cat([kv[0], k], dim=2)
        v = torch.cat([kv[1], v], dim=2)
    new_kv = [k[:, :, -window:, :], v[:, :, -window:, :]] if k.size(2) > window else [k, v]
    attn = torch.softmax(attention_scores(sd, a, q, k, pos_emb, h), dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(x.size(0), -1, d)
    x = _attn_residual(sd, p, x, y, F.linear(o, sd[a + "linear_out.weight"], sd[a + "linear_out.bias"]), concat_after)
    if not normalize_before:
        x = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])                # transformer.py:117-118
    y = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"]) if normalize_before else x
    y, ffn_cache = feed_forward(sd, p, y, kv[2] if (kv is not None and len(kv) > 2) else None)
    if ffn_cache is not None:
        new_kv.append(ffn_cache)                      # third entry of the layer's cache: Conv1dLinear left context
    x = x + y
    if not normalize_before:
        x = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])                # transformer.py:127-128
    return x, new_kv


def layer_offline(sd: State, i: int, x: Tensor, pos_emb: Tensor, mask: Tensor, h: int, normalize_before: bool = True,
                  concat_after: bool = False) -> Tensor:
    """TransformerLayer.forward (transformer.py:75-100) + MultiHeadedAttention.forward
    (attention.py:350-405): masked_fill(min) -> softmax -> masked_fill(0)."""
    p = "enc.1.encoders.%d." % i
    a = p + "self_attn."
    d = x.size(-1)
    y = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"]) if normalize_before else x
    q = _split_heads(F.linear(y, sd[a + "linear_q.weight"], sd[a + "linear_q.bias"]), h)
    k = _split_heads(F.linear(y, sd[a + "linear_k.weight"], sd[a + "linear_k.bias"]), h)
    v = _split_heads(F.linear(y, sd[a + "linear_v.weight"], sd[a + "linear_v.bias"]), h)
    scores = attention_scores(sd, a, q, k, pos_emb, h)
    dead = mask.unsqueeze(1).eq(0)
    attn = torch.softmax(scores.masked_fill(dead, MIN_VALUE), dim=-1).masked_fill(dead, 0.0)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(x.size(0), -1, d)
    x = _attn_residual(sd, p, x, y, F.linear(o, sd[a + "linear_out.weight"], sd[a + "linear_out.bias"]), concat_after)
    if not normalize_before:
        x = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])                # transformer.py:89-90
    y = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"]) if normalize_before else x
    y, _ = feed_forward(sd, p, y)
    x = x + y
    if not normalize_before:
        x = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])                # transformer.py:97-98
    return x


class EncoderOracle:
    """speechEncoder (encoder/encoder.py:45-155) as functions over a state dict.

    cfg needs: d_model, n_heads, n_layers, chunk_size, left_chunks, kv_window, full_chunk_size,
    pe_wrap (``freeze_omni_b200.config.PathConfig`` provides them)."""

    def __init__(self, cfg, sd: State):
        self.cfg, self.sd = cfg, sd

    def new_buffer(self) -> list:
        return [None] * self.cfg.n_layers          # audioLLM.py:377-378

    def infer(self, feats: Tensor, buffer: list, pe_index: int) -> Tuple[Tensor, list, int]:
        """encoder.py:149-155 -> subsampling.py:67-73,103-106 -> transformer.py:267-285.
        feats (B, Tin, F).  Position bookkeeping: attention.py:105-121."""
        c, sd = self.cfg, self.sd
        x = embed(sd, conv_subsample4(sd, cmvn(sd, feats)))
        n_cache = 0 if buffer[0] is None else buffer[0][0].size(2)
        pe_index = pe_index % c.pe_wrap
        x = x * math.sqrt(c.d_model)
        start = max(0, pe_index - c.full_chunk_size)
        pos_emb = pos_table(start, n_cache + x.size(1), c.d_model).unsqueeze(0).to(x.device)   # transformer.py:278-279
        pe_index = pe_index + c.chunk_size
        nb, ca = getattr(c, "normalize_before", True), getattr(c, "concat_after", False)
        for i in range(c.n_layers):
            x, buffer[i] = layer_stream(sd, i, x, pos_emb, buffer[i], c.n_heads, c.kv_window, nb, ca)
        if nb:                                                         # transformer.py:232-233,282-283
            x = F.layer_norm(x, (c.d_model,), sd["enc.1.after_norm.weight"], sd["enc.1.after_norm.bias"])
        return x, buffer, pe_index

    def forward(self, feats: Tensor, ilens: Tensor, chunk: Optional[int] = None,
                left: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """encoder.py:104-147 -> subsampling.py:41-65,98-101 -> transformer.py:237-264 with static
        chunk masks.  Returns (xs (B,T',D), masks (B,1,T'))."""
        c, sd = self.cfg, self.sd
        chunk = c.chunk_size if chunk is None else chunk
        left = c.left_chunks if left is None else left
        t_in = feats.size(1)
        valid = ~pad_mask(ilens, t_in).unsqueeze(1)
        x = conv_subsample4(sd, cmvn(sd, feats))
        valid = valid[:, :, 2::2][:, :, 2::2]                       # subsampling.py:65
        mask = offline_attention_mask(valid, chunk, left)           # transformer.py:253-258
        x = embed(sd, x) * math.sqrt(c.d_model)                     # attention.py:100-102
        pos_emb = pos_table(0, x.size(1), c.d_model).unsqueeze(0).to(x.device)
        nb, ca = getattr(c, "normalize_before", True), getattr(c, "concat_after", False)
        for i in range(c.n_layers):
            x = layer_offline(sd, i, x, pos_emb, mask, c.n_heads, nb, ca)
        if nb:
            x = F.layer_norm(x, (c.d_model,), sd["enc.1.after_norm.weight"], sd["enc.1.after_norm.bias"])
        return x, valid


# ------------------------------------------------------------------------------------------------
# Adapter: CNNSubsampling single-conv branch, models/adapter.py:112-157
# ------------------------------------------------------------------------------------------------
def _bn_eval(y: Tensor, sd: State, name: str) -> Tensor:
    """BatchNorm1d(eps=1e-3) in eval mode over (B, C, T) (adapter.py:25-26,87,92)."""
    return F.batch_norm(y, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.0, 1e-3)


def cnn_adapter_forward(cfg, sd: State, x: Tensor, mask: Tensor) -> Tensor:
    """CNNAdapter (adapter.py:10-57): mask fill, two causal convs (kernel k, stride 1, left zero pad k-1) each with
    BatchNorm + ReLU, then Linear(4C -> E).  No cache: every call starts from a zero left context."""
    k = cfg.adapter_kernel
    xt = x.transpose(1, 2)
    if mask.size(2) > 0:
        xt = xt.masked_fill(~mask, 0.0)                              # adapter.py:41-42
    y = F.relu(_bn_eval(F.conv1d(F.pad(xt, (k - 1, 0)), sd["conv1d1.weight"], sd["conv1d1.bias"]), sd, "bn1"))    # :44-47
    y = F.relu(_bn_eval(F.conv1d(F.pad(y, (k - 1, 0)), sd["conv1d2.weight"], sd["conv1d2.bias"]), sd, "bn2"))     # :49-52
    return F.linear(y.transpose(1, 2), sd["project.weight"], sd["project.bias"])                                  # :54-55


def two_conv_adapter_forward(cfg, sd: State, x: Tensor, mask: Tensor, cache: Optional[List[Optional[Tensor]]] = None):
    """CNNSubsampling with enc_out_dim * 4 < llm_embed_dim (adapter.py:84-96): conv(C -> 2C, k, stride 1) + BN + ReLU, then
    conv(2C -> 4C, k, stride 2) + BN + ReLU, Linear(4C -> E).  cache = [c0 (B, 2C, k-1), c1 (B, C, k-1)] (adapter.py:123-143):
    c1 is the left context of the first conv's INPUT, c0 of the second conv's input."""
    k = cfg.adapter_kernel
    xt = x.transpose(1, 2)
    if mask.size(2) > 0:
        xt = xt.masked_fill(~mask, 0.0)
    if cache is None:
        xt = F.pad(xt, (k - 1, 0))                                   # :124-125
        new_cache: List[Optional[Tensor]] = [None, xt[:, :, 1 - k:].contiguous()]
    else:
        xt = torch.cat((cache[1], xt), dim=2)                        # :126-127
        new_cache = [cache[0], xt[:, :, 1 - k:].contiguous()]
    y = F.relu(_bn_eval(F.conv1d(xt, sd["conv1d1.weight"], sd["conv1d1.bias"]), sd, "bn1"))                       # :132-134
    if new_cache[0] is None:
        y = F.pad(y, (k - 1, 0))                                     # :136-137
    else:
        y = torch.cat((new_cache[0], y), dim=2)                      # :138-139
    new_cache[0] = y[:, :, 1 - k:].contiguous()                      # :140-141
    y = F.relu(_bn_eval(F.conv1d(y, sd["conv1d2.weight"], sd["conv1d2.bias"], stride=2), sd, "bn2"))              # :144-150
    return F.linear(y.transpose(1, 2), sd["project.weight"], sd["project.bias"]), mask[:, :, 0::2], new_cache


def adapter_forward(cfg, sd: State, x: Tensor, mask: Tensor, cache: Optional[List[Tensor]] = None):
    """x (B,T,D), mask (B,1,T) bool, cache None | [Tensor(B,D,k-1)].
    Returns (y (B,T'',E), mask[:, :, 0::2], [new cache])."""
    if getattr(cfg, "adapter_type", "subsampling") == "linear":      # LinearAdapter.forward, adapter.py:69-70: no mask fill, no cache
        return F.linear(x, sd["adpter.weight"], sd["adpter.bias"]), mask, None
    if getattr(cfg, "adapter_type", "subsampling") == "cnn":         # CNNAdapter.forward, adapter.py:33-57 (not built on the GPU yet)
        return cnn_adapter_forward(cfg, sd, x, mask), mask, None
    if cfg.d_model * 4 < cfg.llm_dim:                                # CNNSubsampling two-conv branch, adapter.py:84-96,123-135
        return two_conv_adapter_forward(cfg, sd, x, mask, cache)
    k = cfg.adapter_kernel
    xt = x.transpose(1, 2)
    if mask.size(2) > 0:
        xt = xt.masked_fill(~mask, 0.0)                              # adapter.py:120-121
    if cache is None or cache[0] is None:
        xt = F.pad(xt, (k - 1, 0))                                   # adapter.py:137
    else:
        xt = torch.cat((cache[0], xt), dim=2)                        # adapter.py:139
    new_cache = [xt[:, :, 1 - k:].contiguous()]                      # adapter.py:141,143
    y = F.conv1d(xt, sd["conv1d2.weight"], sd["conv1d2.bias"], stride=2).transpose(1, 2)
    if getattr(cfg, "adapter_norm", "layer") == "batch":             # adapter.py:100-101,146: BatchNorm1d(eps=1e-3), eval mode
        y = F.batch_norm(y.transpose(1, 2), sd["bn2.running_mean"], sd["bn2.running_var"], sd["bn2.weight"], sd["bn2.bias"],
                         False, 0.0, 1e-3).transpose(1, 2)
    else:                                                            # adapter.py:102-103,145-149
        y = F.layer_norm(y, (y.size(-1),), sd["bn2.weight"], sd["bn2.bias"], eps=1e-3)
    y = F.gelu(y) if cfg.adapter_act == "gelu" else F.relu(y)
    y = F.linear(y, sd["project.weight"], sd["project.bias"])
    return y, mask[:, :, 0::2], new_cache


# ------------------------------------------------------------------------------------------------
# Whole-path helpers used by tests and by bench.py's CPU legs
# ------------------------------------------------------------------------------------------------
class StreamSession:
    """One session's state exactly as the service keeps it (bin/dialog_state_pred.py:221-232):
    frontend carry, encoder KV list, adapter cache, pe_index."""

    def __init__(self, cfg, enc_sd: State, adp_sd: State):
        self.cfg, self.enc, self.adp_sd = cfg, EncoderOracle(cfg, enc_sd), adp_sd
        self.front = StreamingFrontend(cfg.sample_rate, cfg.frame_length_ms, cfg.frame_shift_ms,
                                       cfg.frames_per_chunk, cfg.context_frames, cfg.feat_dim)
        self.reset()

    def reset(self):
        self.front.reset()
        self.buffer = self.enc.new_buffer()
        self.cache = None
        self.pe_index = 0

    def step_feats(self, feats: Tensor):
        enc_out, self.buffer, self.pe_index = self.enc.infer(feats, self.buffer, self.pe_index)
        mask = torch.ones(enc_out.shape[0], 1, enc_out.shape[1], dtype=torch.bool)   # audioLLM.py:382
        y, _, self.cache = adapter_forward(self.cfg, self.adp_sd, enc_out, mask, self.cache)
        return enc_out, y

    def step_pcm(self, audio: Tensor, scale: Optional[float] = None):
        feats = self.front.process(audio, self.cfg.pcm_scale if scale is None else scale)
        return (feats,) + self.step_feats(feats)


def offline_path(cfg, enc_sd: State, adp_sd: State, feats: Tensor, ilens: Tensor,
                 chunk: Optional[int] = None, left: Optional[int] = None):
    """speechEncoder.forward then CNNSubsampling.forward(cache=None)."""
    xs, masks = EncoderOracle(cfg, enc_sd).forward(feats, ilens, chunk, left)
    y, ymask, _ = adapter_forward(cfg, adp_sd, xs, masks, None)
    return xs, masks, y, ymask
