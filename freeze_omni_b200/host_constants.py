"""Weights-only constants the host packer hands to the library (fo_load_tensor names "fbank.window",
"fbank.mel", "pos.table").  They are built with the same torch fp32 operations the reference's CPU path
executes, so the device kernels start from bit-identical tables:
  * Povey window and mel filterbank: torchaudio.compliance.kaldi (kaldi.py:98-100, 436-511) as called at
    bin/inference.py:77-78 (dither 0, low_freq 20, high_freq Nyquist, no VTLN);
  * sinusoid table: models/encoder/attention.py:27-34 == the rows RelPositionalEncoding.infer rebuilds on
    the CPU every chunk (attention.py:111-117).
"""
from __future__ import annotations

import math

import torch


def fft_size(frame_len: int) -> int:
    return 1 << (frame_len - 1).bit_length()


def fbank_window(frame_len: int) -> torch.Tensor:
    return torch.hann_window(frame_len, periodic=False, dtype=torch.float32).pow(0.85)


def fbank_mel(n_mel: int, frame_len: int, sample_rate: int, low_freq: float = 20.0) -> torch.Tensor:
    """(n_mel, fft/2 + 1); last column (Nyquist) is the zero pad of kaldi.py:627."""
    padded = fft_size(frame_len)
    nyquist = 0.5 * sample_rate
    n_bins = padded // 2
    width = sample_rate / padded
    lo = 1127.0 * math.log(1.0 + low_freq / 700.0)
    hi = 1127.0 * math.log(1.0 + nyquist / 700.0)
    delta = (hi - lo) / (n_mel + 1)
    idx = torch.arange(n_mel).unsqueeze(1)
    left, center, right = lo + idx * delta, lo + (idx + 1.0) * delta, lo + (idx + 2.0) * delta
    mel = (1127.0 * (1.0 + (width * torch.arange(float(n_bins))) / 700.0).log()).unsqueeze(0)
    tri = torch.max(torch.zeros(1), torch.min((mel - left) / (center - left), (right - mel) / (right - center)))
    return torch.nn.functional.pad(tri, (0, 1)).float().contiguous()


def pos_table(max_len: int, d_model: int) -> torch.Tensor:
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.contiguous()
