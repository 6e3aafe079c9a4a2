"""Checkpoint naming glue for the model-load step of both entry points (ref:scripts/evaluate_model.py:41-75,
ref:scripts/transcribe_single.py:15-33): MLX-named tensors (``decoder.blocks.N.attn.query.weight`` ...) are renamed to the
HF names libwipa's weight table uses; conv weights change from MLX's [out, k, in] to [out, in, k]."""
from __future__ import annotations

import os
import re
from typing import Dict, Mapping

import numpy as np
import torch

from .archs import WhisperArch

_BLOCK = re.compile(r"^(encoder|decoder)\.blocks\.(\d+)\.(.+)$")
_SUB = {
    "attn.query": "self_attn.q_proj", "attn.key": "self_attn.k_proj", "attn.value": "self_attn.v_proj",
    "attn.out": "self_attn.out_proj", "attn_ln": "self_attn_layer_norm",
    "cross_attn.query": "encoder_attn.q_proj", "cross_attn.key": "encoder_attn.k_proj",
    "cross_attn.value": "encoder_attn.v_proj", "cross_attn.out": "encoder_attn.out_proj",
    "cross_attn_ln": "encoder_attn_layer_norm", "mlp1": "fc1", "mlp2": "fc2", "mlp.0": "fc1", "mlp.2": "fc2",
    "mlp_ln": "final_layer_norm",
}
_TOP = {
    "decoder.token_embedding.weight": "model.decoder.embed_tokens.weight",
    "decoder.positional_embedding": "model.decoder.embed_positions.weight",
    "decoder.ln.weight": "model.decoder.layer_norm.weight", "decoder.ln.bias": "model.decoder.layer_norm.bias",
    "encoder.ln_post.weight": "model.encoder.layer_norm.weight", "encoder.ln_post.bias": "model.encoder.layer_norm.bias",
    "encoder.conv1.bias": "model.encoder.conv1.bias", "encoder.conv2.bias": "model.encoder.conv2.bias",
    "encoder.conv1.weight": "model.encoder.conv1.weight", "encoder.conv2.weight": "model.encoder.conv2.weight",
}


def mlx_to_hf_name(name: str) -> str:
    if name.startswith("model.") or name == "proj_out.weight":
        return name
    if name in _TOP:
        return _TOP[name]
    m = _BLOCK.match(name)
    if m:
        side, idx, rest = m.groups()
        stem, _, leaf = rest.rpartition(".")
        if stem in _SUB:
            return f"model.{side}.layers.{idx}.{_SUB[stem]}.{leaf}"
    raise KeyError(f"no HF name for {name!r}")


def to_hf_state_dict(params: Mapping[str, object], arch: WhisperArch) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(torch.float32)
        hf = mlx_to_hf_name(k)
        if k in ("encoder.conv1.weight", "encoder.conv2.weight") and t.dim() == 3 and t.shape[1] == 3:
            t = t.permute(0, 2, 1)                     # MLX conv1d weight is [out, k, in]
        out[hf] = t.contiguous()
    return out


def load_weights_dir(path: str, arch: WhisperArch) -> Dict[str, torch.Tensor]:
    """model.safetensors first, then model.npz / weights.npz, as ref:scripts/evaluate_model.py:41-56."""
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return to_hf_state_dict(load_file(st), arch)
    for name in ("model.npz", "weights.npz"):
        p = os.path.join(path, name)
        if os.path.exists(p):
            return to_hf_state_dict(dict(np.load(p)), arch)
    raise FileNotFoundError(f"no model.safetensors / model.npz under {path!r} (the reference exits here, "
                            "ref:scripts/transcribe_single.py:36-37)")
