"""Drop-in nn.Module surface of the reference's speech encoder and adapter.

Same constructor arguments, attribute names, state-dict keys and call signatures as
``models/encoder/encoder.py:speechEncoder``, ``models/encoder/cmvn.py:GlobalCMVN`` and
``models/adapter.py:CNNSubsampling`` so that ``models/utils.py:init_encoder_llm`` and
``models/audioLLM.py:recognize`` can use them unchanged -- but the modules only HOLD parameters;
every forward/infer call goes through the C ABI into the sm_100a kernels.  No CPU path: calling
them without a CUDA device raises.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch
from torch import nn

from .config import PathConfig, path_config_from_dict
from .engine import Engine

_ENGINES: "weakref.WeakKeyDictionary[nn.Module, Dict[torch.dtype, Engine]]" = weakref.WeakKeyDictionary()


def _compute_dtype(module: nn.Module) -> torch.dtype:
    """bf16 under ``torch.autocast('cuda', torch.bfloat16)`` (models/pipeline.py:67-68), else the
    module's ``compute_dtype`` attribute (default fp32)."""
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return getattr(module, "compute_dtype", torch.float32)


class _Holder(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: compute runs in libfo_b200.so")


class GlobalCMVN(nn.Module):
    """models/encoder/cmvn.py:6-35.  Fused into the first subsampling kernel on the device."""

    def __init__(self, mean: torch.Tensor, istd: torch.Tensor, norm_var: bool = True):
        super().__init__()
        assert mean.shape == istd.shape
        self.norm_var = norm_var          # False: mean removal only (cmvn.py:32-34) -- the engine is then fed istd = 1
        self.register_buffer("mean", mean)
        self.register_buffer("istd", istd)


class LayerCacheView:
    """buffer[i] of the reference is ``[K, V]`` with K,V (1, H, n<=window, d_k) (attention.py:415-428).
    Here it is a lazy view of the session's ring in HBM, materialised only when indexed."""

    def __init__(self, owner: "SessionCache", layer: int):
        self._owner, self._layer = owner, layer

    def __getitem__(self, i):
        ks, vs = [], []
        for sid in self._owner.slots:
            k, v = self._owner.engine.export_kv(int(sid), self._layer)
            ks.append(k)
            vs.append(v)
        return (torch.cat(ks, 0), torch.cat(vs, 0))[i]

    def __len__(self):
        return 2


class SessionCache(list):
    """What ``speechEncoder.infer`` returns in place of the reference's list of per-layer [K, V]: still a
    list of ``num_blocks`` entries (callers only store it and pass it back, bin/dialog_state_pred.py:809-814),
    bound to session slots inside the library.  Slots are released when the object is collected."""

    def __init__(self, engine: Engine, slots, num_blocks: int):
        super().__init__(LayerCacheView(self, i) for i in range(num_blocks))
        self.engine, self.slots = engine, slots
        self.expected_pe = 0
        self._fin = weakref.finalize(self, SessionCache._release, weakref.ref(engine), slots.copy())

    @staticmethod
    def _release(engine_ref, slots):
        eng = engine_ref()
        if eng is not None and getattr(eng, "_h", None):
            try:
                eng.free(slots)
            except Exception:
                pass

    def export_reference(self) -> List[List[torch.Tensor]]:
        return [[self[i][0], self[i][1]] for i in range(len(self))]


class speechEncoder(nn.Module):
    """models/encoder/encoder.py:45-155."""

    def __init__(self, input_dim, overview_conf=None, para_conf=None, global_cmvn=None):
        super().__init__()
        cfg = path_config_from_dict({"input_dim": input_dim,
                                     "encoder_conf": {"overview_conf": overview_conf or {}, "para_conf": para_conf or {}}},
                                    encoder_only=True)
        self.path_config: PathConfig = cfg
        self.config = ["subsampling", "transformer"]
        self.global_cmvn = global_cmvn
        d, f2 = cfg.d_model, cfg.sub_freq
        sub = _Holder()
        sub.core = _Holder()
        sub.core.conv = nn.Sequential(nn.Conv2d(1, d, 3, 2), nn.ReLU(), nn.Conv2d(d, d, 3, 2), nn.ReLU())
        sub.core.out = nn.Sequential(nn.Linear(d * f2, d))
        sub.subsampling_rate = 4
        tr = _Holder()
        if cfg.input_layer == "linear":
            tr.embed = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d), nn.Dropout(0.1), nn.ReLU())
        else:
            tr.embed = nn.Sequential(nn.Identity())
        layers = []
        for _ in range(cfg.n_layers):
            lay = _Holder()
            att = _Holder()
            att.linear_q, att.linear_k = nn.Linear(d, d), nn.Linear(d, d)
            att.linear_v, att.linear_out = nn.Linear(d, d), nn.Linear(d, d)
            att.linear_pos = nn.Linear(d, d, bias=False)
            att.pos_bias_u = nn.Parameter(torch.empty(cfg.n_heads, cfg.d_k))
            att.pos_bias_v = nn.Parameter(torch.empty(cfg.n_heads, cfg.d_k))
            nn.init.xavier_uniform_(att.pos_bias_u)
            nn.init.xavier_uniform_(att.pos_bias_v)
            ff = _Holder()
            if cfg.ffn_type == "conv1d-linear":        # Conv1dLinear (attention.py:217-234)
                ff.w_1 = nn.Sequential(nn.Conv1d(d, d, cfg.ffn_conv_kernel, groups=d), nn.Conv1d(d, cfg.ffn_dim, 1))
            elif cfg.ffn_type == "conv1d":             # MultiLayeredConv1d (attention.py:158-184)
                kf = cfg.ffn_conv_kernel
                ff.w_1 = nn.Conv1d(d, cfg.ffn_dim, kf, stride=1, padding=(kf - 1) // 2)
            else:
                ff.w_1 = nn.Linear(d, cfg.ffn_dim)
            if cfg.ffn_type == "conv1d":
                ff.w_2 = nn.Conv1d(cfg.ffn_dim, d, cfg.ffn_conv_kernel, stride=1, padding=(cfg.ffn_conv_kernel - 1) // 2)
            else:
                ff.w_2 = nn.Linear(cfg.ffn_dim, d)
            lay.self_attn, lay.feed_forward = att, ff
            lay.norm1, lay.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)
            if cfg.concat_after:                       # transformer.py:69-70
                lay.concat_linear = nn.Linear(2 * d, d)
            layers.append(lay)
        tr.encoders = nn.ModuleList(layers)
        if cfg.normalize_before:                       # transformer.py:232-233
            tr.after_norm = nn.LayerNorm(d)
        # attributes callers read or set (audioLLM.py:378; encoder.py:132-138)
        tr.num_blocks = cfg.n_layers
        tr.chunk_size, tr.left_chunks = cfg.chunk_size, cfg.left_chunks
        tr.transformer_dynamic_chunks = cfg.dynamic_chunks
        self.enc = nn.ModuleList([sub, tr])
        self._output_size = d
        self.max_sessions = 64
        self.compute_dtype = torch.float32
        num_params = sum(p.numel() for p in self.parameters())
        print('the number of speech encoder params: {}M'.format(num_params / 1024 / 1024))

    def output_size(self) -> int:
        return self._output_size

    # ---- engine management --------------------------------------------------------------------
    def invalidate(self) -> None:
        """Drop device copies of the weights (call after changing parameters)."""
        for eng in _ENGINES.pop(self, {}).values():
            eng.close()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.invalidate()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _apply(self, fn, recurse=True):
        self.invalidate()
        return super()._apply(fn, recurse)

    def engine(self, dtype: Optional[torch.dtype] = None) -> Engine:
        dtype = dtype or _compute_dtype(self)
        engines = _ENGINES.setdefault(self, {})
        if dtype not in engines:
            p = next(self.parameters())
            if p.device.type != "cuda":
                raise RuntimeError("speechEncoder: move the module to a CUDA device first; there is no CPU path")
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            if self.global_cmvn is not None and not self.global_cmvn.norm_var:
                sd["global_cmvn.istd"] = torch.ones_like(sd["global_cmvn.istd"])      # x - mean only (cmvn.py:32-34)
            engines[dtype] = Engine(self.path_config, enc_state=sd, adp_state=None, dtype=dtype,
                                    device=p.device.index or 0, max_sessions=self.max_sessions,
                                    use_cmvn=self.global_cmvn is not None)
        return engines[dtype]

    # ---- encoder.py:104-147 -------------------------------------------------------------------
    @torch.compiler.disable
    @torch.no_grad()
    def forward(self, xs, ilens, decoding_chunk_size=None, num_decoding_left_chunks=None):
        tr = self.enc[1]
        if decoding_chunk_size is not None and num_decoding_left_chunks is not None:
            tr.chunk_size, tr.left_chunks = decoding_chunk_size, num_decoding_left_chunks
            tr.transformer_dynamic_chunks = False
        if tr.transformer_dynamic_chunks:
            raise NotImplementedError("transformer-dynamic-chunks draws a random chunk size even in eval "
                                      "(transformer.py:245-251); pass decoding_chunk_size/num_decoding_left_chunks")
        assert xs.dim() == 3
        xs = xs.float().contiguous()
        enc, mask, _, _ = self.engine().encode_offline(xs, ilens, tr.chunk_size, tr.left_chunks, want_adapter=False)
        return enc, mask

    # ---- encoder.py:149-155 -------------------------------------------------------------------
    @torch.compiler.disable
    @torch.no_grad()
    def infer(self, xs_pad, buffer, buffer_index, buffer_out, pe_index):
        eng = self.engine()
        L = self.path_config.n_layers
        if isinstance(buffer, SessionCache):
            cache = buffer
            if cache.engine is not eng:
                raise RuntimeError("encoder cache was created under a different compute dtype")
        elif isinstance(buffer, list) and len(buffer) == L and all(b is None for b in buffer):
            cache = SessionCache(eng, eng.alloc(int(xs_pad.size(0))), L)          # audioLLM.py:377-378
        else:
            raise TypeError("buffer must be [None]*num_blocks or the cache returned by a previous infer(); "
                            "use import_reference_cache() to adopt a reference-layout cache")
        if len(cache.slots) != xs_pad.size(0):
            raise ValueError("batch size changed between streaming calls")
        pe_index = int(pe_index)
        if pe_index != cache.expected_pe:
            eng.set_pe_index(cache.slots, pe_index)
        xs = xs_pad.float().contiguous()
        enc_out, _ = eng.encode_stream(cache.slots, xs, want_adapter=False)
        pe_next = pe_index % self.path_config.pe_wrap + self.path_config.chunk_size   # attention.py:107,120
        cache.expected_pe = pe_next
        return enc_out, cache, buffer_index + L, buffer_out, pe_next

    def import_reference_cache(self, buffer: List[List[torch.Tensor]], pe_index: int,
                               frames_seen: Optional[int] = None) -> SessionCache:
        """Adopt a cache in the reference layout (list of [K, V], (1, H, n, d_k)) for one session."""
        eng = self.engine()
        cache = SessionCache(eng, eng.alloc(1), self.path_config.n_layers)
        n = int(buffer[0][0].size(2)) if buffer[0] is not None else 0
        eng.set_frames(int(cache.slots[0]), n if frames_seen is None else frames_seen)
        for i, kv in enumerate(buffer):
            if kv is not None:
                eng.import_kv(int(cache.slots[0]), i, kv[0], kv[1])
        eng.set_pe_index(cache.slots, pe_index)
        cache.expected_pe = int(pe_index)
        return cache


class _AdapterBase(nn.Module):
    """What the adapter drop-ins share: one Engine per compute dtype, rebuilt when the parameters change."""

    def invalidate(self) -> None:
        for eng in _ENGINES.pop(self, {}).values():
            eng.close()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.invalidate()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _apply(self, fn, recurse=True):
        self.invalidate()
        return super()._apply(fn, recurse)

    def engine(self, dtype: Optional[torch.dtype] = None) -> Engine:
        dtype = dtype or _compute_dtype(self)
        engines = _ENGINES.setdefault(self, {})
        if dtype not in engines:
            p = next(self.parameters())
            if p.device.type != "cuda":
                raise RuntimeError("%s: move the module to a CUDA device first; there is no CPU path" % type(self).__name__)
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            engines[dtype] = Engine(self.path_config, enc_state=None, adp_state=sd, dtype=dtype,
                                    device=p.device.index or 0, max_sessions=1)
        return engines[dtype]


class CNNSubsampling(_AdapterBase):
    """models/adapter.py:72-157: the single-conv branch (enc_out_dim * 4 >= llm_embed_dim) and the two-conv branch
    (:84-96; two BatchNorm1d + ReLU convolutions, two caches)."""

    def __init__(self, enc_out_dim: int = 512, llm_embed_dim: int = 4096, kernel_size: int = 5,
                 activation_func: str = 'relu', norm: str = 'batch'):
        super().__init__()
        self.kernel_size = kernel_size
        if enc_out_dim * 4 < llm_embed_dim:
            self.left_padding1 = nn.ConstantPad1d((kernel_size - 1, 0), 0.0)
            self.conv1d1 = nn.Conv1d(enc_out_dim, 2 * enc_out_dim, kernel_size, 1, 0)
            self.bn1 = nn.BatchNorm1d(2 * enc_out_dim, eps=1e-3, momentum=0.99)
            self.relu1 = nn.ReLU()
            self.left_padding2 = nn.ConstantPad1d((kernel_size - 1, 0), 0.0)
            self.conv1d2 = nn.Conv1d(2 * enc_out_dim, 4 * enc_out_dim, kernel_size, 2, 0)
            self.bn2 = nn.BatchNorm1d(4 * enc_out_dim, eps=1e-3, momentum=0.99)
            self.relu2 = nn.ReLU()
            self.project = nn.Linear(4 * enc_out_dim, llm_embed_dim)
            self.cnn_num = 2
        else:
            if norm not in ('layer', 'batch'):
                raise NotImplementedError("adapter norm must be 'layer' or 'batch' (adapter.py:100-103)")
            self.left_padding2 = nn.ConstantPad1d((kernel_size - 1, 0), 0.0)
            self.conv1d2 = nn.Conv1d(enc_out_dim, 2 * enc_out_dim, kernel_size, 2, 0)
            self.bn2 = (nn.LayerNorm(2 * enc_out_dim, eps=1e-3) if norm == 'layer'
                        else nn.BatchNorm1d(2 * enc_out_dim, eps=1e-3, momentum=0.99))
            self.relu2 = nn.GELU() if activation_func == 'gelu' else nn.ReLU()
            self.project = nn.Linear(2 * enc_out_dim, llm_embed_dim)
            self.cnn_num = 1
        self.path_config = PathConfig(d_model=enc_out_dim, n_heads=enc_out_dim // 64, llm_dim=llm_embed_dim,
                                      adapter_kernel=kernel_size,
                                      adapter_act='gelu' if activation_func == 'gelu' else 'relu',
                                      adapter_norm=norm if norm in ('layer', 'batch') else 'batch')
        self.compute_dtype = torch.float32

    @torch.compiler.disable
    @torch.no_grad()
    def forward(self, x, mask_pad, cache=None, return_cache=False):
        m = mask_pad if mask_pad.size(2) > 0 else None
        if self.cnn_num == 2:
            y, new_cache = self.engine().adapter_forward(x, m, cache)
        else:
            y, nc = self.engine().adapter_forward(x, m, None if cache is None else cache[0])
            new_cache = [nc]
        if return_cache:
            return y, mask_pad[:, :, 0::2], new_cache
        return y, mask_pad[:, :, 0::2]


class CNNAdapter(_AdapterBase):
    """models/adapter.py:10-57 (adpter_type == 'cnn', the AudioLLM default, audioLLM.py:159-160): two causal stride-1
    convolutions with BatchNorm1d + ReLU and Linear(4 * enc_out_dim -> llm_embed_dim); no cache, frame rate kept."""

    def __init__(self, enc_out_dim: int = 512, llm_embed_dim: int = 4096, kernel_size: int = 5):
        super().__init__()
        self.left_padding1 = nn.ConstantPad1d((kernel_size - 1, 0), 0.0)
        self.left_padding2 = nn.ConstantPad1d((kernel_size - 1, 0), 0.0)
        self.conv1d1 = nn.Conv1d(enc_out_dim, 2 * enc_out_dim, kernel_size, 1, 0)
        self.conv1d2 = nn.Conv1d(2 * enc_out_dim, 4 * enc_out_dim, kernel_size, 1, 0)
        self.bn1 = nn.BatchNorm1d(2 * enc_out_dim, eps=1e-3, momentum=0.99)
        self.bn2 = nn.BatchNorm1d(4 * enc_out_dim, eps=1e-3, momentum=0.99)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()
        self.project = nn.Linear(4 * enc_out_dim, llm_embed_dim)
        self.path_config = PathConfig(d_model=enc_out_dim, n_heads=max(1, enc_out_dim // 64), llm_dim=llm_embed_dim,
                                      adapter_kernel=kernel_size, adapter_act='relu', adapter_norm='batch', adapter_type='cnn')
        self.compute_dtype = torch.float32

    @torch.compiler.disable
    @torch.no_grad()
    def forward(self, x, mask_pad):
        y, _ = self.engine().adapter_forward(x, mask_pad if mask_pad.size(2) > 0 else None, None)
        return y, mask_pad


class LinearAdapter(_AdapterBase):
    """models/adapter.py:59-70 (adpter_type == 'linear', audioLLM.py:161-162): y = Linear(enc_out_dim -> llm_embed_dim)(x),
    mask passed through; same state-dict keys (adpter.weight / adpter.bias)."""

    def __init__(self, enc_out_dim: int = 512, llm_embed_dim: int = 4096):
        super().__init__()
        self.adpter = nn.Linear(enc_out_dim, llm_embed_dim)
        self.path_config = PathConfig(d_model=enc_out_dim, n_heads=max(1, enc_out_dim // 64), llm_dim=llm_embed_dim,
                                      adapter_type='linear')
        self.compute_dtype = torch.float32

    @torch.compiler.disable
    @torch.no_grad()
    def forward(self, x, mask_pad):
        y, _ = self.engine().adapter_forward(x, None, None)
        return y, mask_pad
