"""Seeded random-init weights under the reference's state-dict key names.

The real ``checkpoints/audiollm/final.pt`` is not in the reference tree, so every
parity case runs on random-init weights.  Rather than depending on the reference modules'
constructor RNG order, the tensors are a pure function of (config, seed, key name); the golden
generator loads this state dict INTO the reference modules, the oracle and the CUDA path read it
directly.  Key names/shapes are those of SURVEY 3.4 (``models/encoder/*``, ``models/adapter.py``).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import torch

from .config import PathConfig


def encoder_param_shapes(cfg: PathConfig) -> List[Tuple[str, Tuple[int, ...], int]]:
    """(key, shape, fan_in) of every encoder tensor, in module order (subsampling.py:26-34,
    transformer.py:186-234, attention.py:283-307)."""
    d, f, ff = cfg.d_model, cfg.feat_dim, cfg.ffn_dim
    out: List[Tuple[str, Tuple[int, ...], int]] = [
        ("global_cmvn.mean", (f,), 0), ("global_cmvn.istd", (f,), 0),
        ("enc.0.core.conv.0.weight", (d, 1, 3, 3), 9), ("enc.0.core.conv.0.bias", (d,), 9),
        ("enc.0.core.conv.2.weight", (d, d, 3, 3), 9 * d), ("enc.0.core.conv.2.bias", (d,), 9 * d),
        ("enc.0.core.out.0.weight", (d, d * cfg.sub_freq), d * cfg.sub_freq),
        ("enc.0.core.out.0.bias", (d,), d * cfg.sub_freq),
    ]
    if cfg.input_layer == "linear":
        out += [("enc.1.embed.0.weight", (d, d), d), ("enc.1.embed.0.bias", (d,), d),
                ("enc.1.embed.1.weight", (d,), -1), ("enc.1.embed.1.bias", (d,), -2)]
    for i in range(cfg.n_layers):
        p = "enc.1.encoders.%d." % i
        for nm in ("linear_q", "linear_k", "linear_v", "linear_out"):
            out += [(p + "self_attn.%s.weight" % nm, (d, d), d), (p + "self_attn.%s.bias" % nm, (d,), d)]
        out += [(p + "self_attn.linear_pos.weight", (d, d), d),
                (p + "self_attn.pos_bias_u", (cfg.n_heads, cfg.d_k), -3),
                (p + "self_attn.pos_bias_v", (cfg.n_heads, cfg.d_k), -3),
                ]
        if cfg.ffn_type == "conv1d-linear":        # Conv1dLinear (attention.py:217-233): depthwise (d,1,k) + pointwise (ff,d,1)
            kf = cfg.ffn_conv_kernel
            out += [(p + "feed_forward.w_1.0.weight", (d, 1, kf), kf), (p + "feed_forward.w_1.0.bias", (d,), kf),
                    (p + "feed_forward.w_1.1.weight", (ff, d, 1), d), (p + "feed_forward.w_1.1.bias", (ff,), d)]
        elif cfg.ffn_type == "conv1d":             # MultiLayeredConv1d (attention.py:158-184): Conv1d(d, ff, k), Conv1d(ff, d, k)
            kf = cfg.ffn_conv_kernel
            out += [(p + "feed_forward.w_1.weight", (ff, d, kf), d * kf), (p + "feed_forward.w_1.bias", (ff,), d * kf)]
        else:
            out += [(p + "feed_forward.w_1.weight", (ff, d), d), (p + "feed_forward.w_1.bias", (ff,), d)]
        if getattr(cfg, "concat_after", False):            # transformer.py:69-70: Linear(size + size, size)
            out += [(p + "concat_linear.weight", (d, 2 * d), 2 * d), (p + "concat_linear.bias", (d,), 2 * d)]
        if cfg.ffn_type == "conv1d":
            out += [(p + "feed_forward.w_2.weight", (d, ff, cfg.ffn_conv_kernel), ff * cfg.ffn_conv_kernel),
                    (p + "feed_forward.w_2.bias", (d,), ff * cfg.ffn_conv_kernel)]
        else:
            out += [(p + "feed_forward.w_2.weight", (d, ff), ff), (p + "feed_forward.w_2.bias", (d,), ff)]
        out += [
                (p + "norm1.weight", (d,), -1), (p + "norm1.bias", (d,), -2),
                (p + "norm2.weight", (d,), -1), (p + "norm2.bias", (d,), -2)]
    if getattr(cfg, "normalize_before", True):             # transformer.py:232-233
        out += [("enc.1.after_norm.weight", (d,), -1), ("enc.1.after_norm.bias", (d,), -2)]
    return out


def adapter_param_shapes(cfg: PathConfig) -> List[Tuple[str, Tuple[int, ...], int]]:
    """CNNSubsampling single-conv branch (adapter.py:97-110)."""
    d, k, e = cfg.d_model, cfg.adapter_kernel, cfg.llm_dim
    if getattr(cfg, "adapter_type", "subsampling") == "linear":      # LinearAdapter (adapter.py:59-70): self.adpter = Linear
        return [("adpter.weight", (e, d), d), ("adpter.bias", (e,), d)]
    if getattr(cfg, "adapter_type", "subsampling") == "cnn" or d * 4 < e:
        # CNNAdapter (adapter.py:10-31) and the two-conv CNNSubsampling branch (adapter.py:84-96) share their tensors
        return [("conv1d1.weight", (2 * d, d, k), d * k), ("conv1d1.bias", (2 * d,), d * k),
                ("bn1.weight", (2 * d,), -1), ("bn1.bias", (2 * d,), -2), ("bn1.running_mean", (2 * d,), -2), ("bn1.running_var", (2 * d,), -4),
                ("conv1d2.weight", (4 * d, 2 * d, k), 2 * d * k), ("conv1d2.bias", (4 * d,), 2 * d * k),
                ("bn2.weight", (4 * d,), -1), ("bn2.bias", (4 * d,), -2), ("bn2.running_mean", (4 * d,), -2), ("bn2.running_var", (4 * d,), -4),
                ("project.weight", (e, 4 * d), 4 * d), ("project.bias", (e,), 4 * d)]
    out = [("conv1d2.weight", (2 * d, d, k), d * k), ("conv1d2.bias", (2 * d,), d * k),
           ("bn2.weight", (2 * d,), -1), ("bn2.bias", (2 * d,), -2)]
    if cfg.adapter_norm == "batch":            # BatchNorm1d buffers (eval mode uses the running statistics)
        out += [("bn2.running_mean", (2 * d,), -2), ("bn2.running_var", (2 * d,), -4)]
    return out + [("project.weight", (e, 2 * d), 2 * d), ("project.bias", (e,), 2 * d)]


def _fill(key: str, shape: Tuple[int, ...], fan_in: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    if key.endswith("global_cmvn.mean"):
        return 9.0 + 2.0 * torch.randn(shape, generator=g)
    if key.endswith("global_cmvn.istd"):
        return 1.0 / (3.0 + torch.rand(shape, generator=g))
    if fan_in == -1:      # LayerNorm gain
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if fan_in == -2:      # LayerNorm shift
        return 0.1 * torch.randn(shape, generator=g)
    if fan_in == -4:      # BatchNorm running variance: positive
        return 0.5 + torch.rand(shape, generator=g)
    if fan_in == -3:      # pos_bias_{u,v}: xavier-uniform bound (attention.py:306-307)
        b = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    b = 1.0 / math.sqrt(fan_in)   # nn.Linear / nn.Conv default bound
    return (torch.rand(shape, generator=g) * 2 - 1) * b


def make_encoder_state(cfg: PathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: _fill(k, s, f, seed).float().contiguous() for k, s, f in encoder_param_shapes(cfg)}


def make_adapter_state(cfg: PathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: _fill("adapter." + k, s, f, seed).float().contiguous() for k, s, f in adapter_param_shapes(cfg)}


def audit_state_dict(expected: List[Tuple[str, Tuple[int, ...], int]], sd: Dict[str, torch.Tensor],
                     reject_unexpected: bool = False) -> None:
    """``load_state_dict(strict=False)`` (models/utils.py:20) silently drops renamed keys; the
    drop-in refuses to run on a partial or mis-shaped state dict instead.  With ``reject_unexpected`` (checkpoint ingest)
    a tensor the configured variant does not own is an error too: ``bn2.running_mean`` under norm='layer' or
    ``conv1d1.*`` under the single-conv adapter mean the yaml describes a different module than the checkpoint holds."""
    missing = [k for k, _, _ in expected if k not in sd]
    if missing:
        raise KeyError("state dict is missing %d tensors, e.g. %s" % (len(missing), missing[:3]))
    for k, shp, _ in expected:
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError("tensor %s has shape %s, expected %s" % (k, tuple(sd[k].shape), shp))
    if reject_unexpected:
        names = {k for k, _, _ in expected}
        has_bn = {k.rsplit(".", 1)[0] for k in names if k.endswith(".running_mean")}
        extra = [k for k in sd if k not in names
                 and not (k.endswith(".num_batches_tracked") and k.rsplit(".", 1)[0] in has_bn)]
        if extra:
            raise ValueError("state dict holds %d tensors the configured variant does not use, e.g. %s -- the yaml's "
                             "model_conf / encoder_conf does not describe this checkpoint" % (len(extra), sorted(extra)[:3]))
