"""CPU-side checks of the boundary: the library builds, loads, exports every symbol include/fo_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from freeze_omni_b200 import build
    build.build()
    from freeze_omni_b200 import _lib
    return _lib.load()


def test_header_symbols_all_exported_and_bound(lib):
    from freeze_omni_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fo_b200.h")).read()
    declared = set(re.findall(r"\b(fo_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.fo_abi_version() == 6


def test_struct_layout_matches_header():
    from freeze_omni_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fo_b200.h")).read()
    body = hdr[hdr.index("typedef struct fo_config {"):hdr.index("} fo_config;")]
    fields = re.findall(r"int32_t\s+([a-z_0-9]+);", body)
    assert fields == [f for f, _ in _lib.FoConfig._fields_]
    body = hdr[hdr.index("typedef struct fo_stats_t {"):hdr.index("} fo_stats_t;")]
    fields = re.findall(r"int64_t\s+([a-z_0-9]+);", body)
    assert fields == [f for f, _ in _lib.FoStats._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from freeze_omni_b200 import _lib
    from freeze_omni_b200.config import load_path_config
    from freeze_omni_b200.engine import Engine
    from freeze_omni_b200.weights import make_adapter_state, make_encoder_state
    h = C.c_void_p()
    cfg = _lib.FoConfig(feat_dim=80, d_model=128, n_heads=2, ffn_dim=256, n_layers=1, chunk_size=4, left_chunks=16,
                        input_layer_linear=1, pos_max_len=5000, llm_dim=256, adapter_kernel=5, adapter_gelu=1,
                        has_encoder=1, has_adapter=1, sample_rate=16000, frame_len=400, frame_shift=160,
                        frames_per_chunk=16, context_frames=3, max_sessions=2, max_stream_frames=19)
    assert lib.fo_create(C.byref(cfg), 0, 0, C.byref(h)) != 0
    assert b"no CUDA device" in lib.fo_last_error()
    c = load_path_config("tiny")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(c, make_encoder_state(c, 3), make_adapter_state(c, 3))


def test_host_constants_match_torchaudio():
    kaldi = pytest.importorskip("torchaudio.compliance.kaldi")
    from freeze_omni_b200 import host_constants as hc
    w = kaldi._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32)
    assert torch.equal(hc.fbank_window(400), w)
    mel = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0]
    assert torch.equal(hc.fbank_mel(80, 400, 16000)[:, :256], mel)
    assert float(hc.fbank_mel(80, 400, 16000)[:, 256].abs().max()) == 0.0
    mel = kaldi.get_mel_banks(80, 256, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0]
    assert torch.equal(hc.fbank_mel(80, 256, 16000)[:, :128], mel)


def test_pos_table_matches_streaming_rows():
    """Rows of the host-built table equal what RelPositionalEncoding.infer computes per chunk
    (attention.py:111-117), restated by the oracle."""
    from freeze_omni_b200 import host_constants as hc
    from oracle import freeze_omni_oracle as O
    tab = hc.pos_table(5000, 128)
    for start, n in ((0, 4), (0, 68), (100, 68), (4860, 72)):
        assert torch.equal(tab[start:start + n], O.pos_table(start, n, 128))


def test_dropin_state_dict_keys_match_reference_names():
    from freeze_omni_b200.config import load_path_config, load_yaml
    from freeze_omni_b200.modules import CNNSubsampling, GlobalCMVN, speechEncoder
    from freeze_omni_b200.weights import adapter_param_shapes, encoder_param_shapes
    y, cfg = load_yaml("tiny"), load_path_config("tiny")
    enc = speechEncoder(80, global_cmvn=GlobalCMVN(torch.zeros(80), torch.ones(80)), **y["encoder_conf"])
    want = {k: tuple(s) for k, s, _ in encoder_param_shapes(cfg)}
    got = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert got == want
    mc = y["model_conf"]
    adp = CNNSubsampling(mc["enc_out_dim"], mc["llm_embed_dim"], mc["kernel_size"], mc["activation_func"], mc["norm"])
    assert {k: tuple(v.shape) for k, v in adp.state_dict().items()} == {k: tuple(s) for k, s, _ in adapter_param_shapes(cfg)}
    assert enc.enc[1].num_blocks == cfg.n_layers and enc.output_size() == cfg.d_model
    import copy
    copy.deepcopy(enc)          # audioLLM.py:68 deep-copies the encoder
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc.infer(torch.zeros(1, 19, 80), [None] * cfg.n_layers, 0, None, 0)
