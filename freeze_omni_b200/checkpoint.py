"""Trained-weights ingest: CMVN statistics files, `final.pt` / `train.yaml`.

Restates models/encoder/cmvn.py:37-107 (load_cmvn and its two parsers) and the loading half of
models/utils.py:11-49 (load_checkpoint / init_encoder_llm) for the drop-in: the checkpoint of AudioLLM holds the
encoder twice (`encoder_user.*`, `encoder_system.*`: audioLLM.py:67-68) and the adapter twice (`adpter_user.*`,
`adpter_system.*`: :160-166); older checkpoints use `encoder.*` / `adpter.*`.  Where the reference loads with
strict=False (utils.py:20) and silently drops anything renamed, this loader audits every expected key and shape.
"""
import json
import os
import re
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from .config import PathConfig, load_yaml, path_config_from_dict
from .weights import adapter_param_shapes, audit_state_dict, encoder_param_shapes


def _finish(sum_stat: np.ndarray, sq_stat: np.ndarray, count: float) -> Tuple[np.ndarray, np.ndarray]:
    """mean = sum/count;  istd = 1/sqrt(max(sq/count - mean^2, 1e-20))  in float64, the reference's arithmetic
    (cmvn.py:52-58, 97-103)."""
    mean = np.asarray(sum_stat, np.float64) / count
    var = np.asarray(sq_stat, np.float64) / count - mean * mean
    var = np.where(var < 1.0e-20, 1.0e-20, var)
    return mean, 1.0 / np.sqrt(var)


def load_cmvn(cmvn_file: str, is_json: bool) -> Tuple[np.ndarray, np.ndarray]:
    """(mean, istd) float64 arrays of length feat_dim; feed them to GlobalCMVN(torch.from_numpy(.).float(), ...)
    exactly as models/utils.py:38-45 does."""
    if is_json:                                              # cmvn.py:37-59
        with open(cmvn_file) as f:
            st = json.load(f)
        return _finish(st["mean_stat"], st["var_stat"], st["frame_num"])
    with open(cmvn_file, "rb") as f:                         # cmvn.py:61-104: kaldi TEXT matrix "[ sums count \n sqs 0 ]"
        raw = f.read()
    if raw[:2] == b"\0B":
        raise ValueError("kaldi binary cmvn is not supported; recompute with compute-cmvn-stats --binary=false")
    arr = raw.decode().split()
    if len(arr) < 6 or arr[0] != "[" or arr[-1] != "]" or arr[-2] != "0" or (len(arr) - 4) % 2:
        raise ValueError("%s is not a 2-row kaldi text cmvn matrix" % cmvn_file)
    d = (len(arr) - 4) // 2
    sums = np.array([float(v) for v in arr[1:d + 1]])
    count = float(arr[d + 1])
    sqs = np.array([float(v) for v in arr[d + 2:2 * d + 2]])
    return _finish(sums, sqs, count)


def split_checkpoint(ckpt: Dict[str, torch.Tensor], role: str = "user") -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(encoder state, adapter state) under the key names of speechEncoder / CNNSubsampling for one role."""
    def pick(prefixes):
        for p in prefixes:
            sub = {k[len(p):]: v for k, v in ckpt.items() if k.startswith(p)}
            if sub:
                return sub
        return {}
    enc = pick(["encoder_%s." % role, "encoder."])
    adp = pick(["adpter_%s." % role, "adpter."])
    # torch.compile wrappers prefix parameters with _orig_mod. (audioLLM.py:266-287)
    enc = {re.sub(r"^_orig_mod\.", "", k): v for k, v in enc.items()}
    adp = {re.sub(r"^_orig_mod\.", "", k): v for k, v in adp.items()}
    return enc, adp


def load_trained(model_dir_or_pt: str, yaml_path: Optional[str] = None, role: str = "user",
                 cmvn_file: Optional[str] = None, is_json_cmvn: Optional[bool] = None):
    """Read `<dir>/final.pt` + `<dir>/train.yaml` (or explicit paths) -> (PathConfig, encoder state, adapter state),
    every expected tensor present and shaped; CMVN from the yaml's cmvn_file unless the checkpoint carries it."""
    pt = model_dir_or_pt if model_dir_or_pt.endswith(".pt") else os.path.join(model_dir_or_pt, "final.pt")
    if yaml_path is None:
        cand = [re.sub(r"\.pt$", ".yaml", pt), os.path.join(os.path.dirname(pt), "train.yaml")]
        yaml_path = next((c for c in cand if os.path.exists(c)), None)
        if yaml_path is None:
            raise FileNotFoundError("no yaml next to %s (tried %s)" % (pt, cand))
    configs: Dict[str, Any] = load_yaml(yaml_path)
    cfg: PathConfig = path_config_from_dict(configs)
    ckpt = torch.load(pt, map_location="cpu", weights_only=True)
    enc, adp = split_checkpoint(ckpt, role)
    cmvn_file = cmvn_file if cmvn_file is not None else configs.get("cmvn_file")
    if "global_cmvn.mean" not in enc and cmvn_file:
        is_json = bool(configs.get("is_json_cmvn", False)) if is_json_cmvn is None else is_json_cmvn
        mean, istd = load_cmvn(cmvn_file, is_json)
        enc["global_cmvn.mean"] = torch.from_numpy(mean).float()
        enc["global_cmvn.istd"] = torch.from_numpy(istd).float()
    exp = encoder_param_shapes(cfg)
    if "global_cmvn.mean" not in enc:
        exp = [e for e in exp if not e[0].startswith("global_cmvn.")]
    audit_state_dict(exp, enc, reject_unexpected=True)
    audit_state_dict(adapter_param_shapes(cfg), adp, reject_unexpected=True)
    return cfg, enc, adp
