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


# Entries of an MLX checkpoint that are not weights of the network: mlx_whisper's public `alignment_heads` array is part of
# `model.parameters()` and therefore of what the reference saves (ref:scripts/train_whisper_ipa.py:420-422); the encoder's
# sinusoidal position table is a private constant there (`_positional_embedding`) and is rebuilt by libwipa at context
# creation (ctx.cu) when no tensor supplies it.
_NOT_WEIGHTS = ("alignment_heads", "encoder.positional_embedding", "encoder._positional_embedding")


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
    """MLX- or HF-named tensors -> HF-named fp32 tensors.  Non-weight entries (`alignment_heads`, the encoder's position
    constant) are skipped; a name that is neither raises KeyError."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in params.items():
        if k in _NOT_WEIGHTS:
            continue
        t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(torch.float32)
        hf = mlx_to_hf_name(k)
        if k in ("encoder.conv1.weight", "encoder.conv2.weight") and t.dim() == 3 and t.shape[1] == 3:
            t = t.permute(0, 2, 1)                     # MLX conv1d weight is [out, k, in]
        out[hf] = t.contiguous()
    return out


_HF_TO_MLX_SUB = {v: k for k, v in _SUB.items() if not k.startswith("mlp.")}
_HF_TO_MLX_TOP = {v: k for k, v in _TOP.items()}
_HF_LAYER = re.compile(r"^model\.(encoder|decoder)\.layers\.(\d+)\.(.+)\.(weight|bias)$")


def hf_to_mlx_state_dict(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The inverse naming / layout map: an HF state dict -> the flat key set mlx_whisper's ``model.parameters()`` has (what
    the reference's checkpoints contain, ref:scripts/train_whisper_ipa.py:420-422).  The tied ``proj_out.weight`` and the
    encoder's sinusoid table have no MLX entry."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in state_dict.items():
        if k in ("proj_out.weight", "model.encoder.embed_positions.weight"):
            continue
        t = torch.as_tensor(v)
        if k in _HF_TO_MLX_TOP:
            name = _HF_TO_MLX_TOP[k]
            if name in ("encoder.conv1.weight", "encoder.conv2.weight"):
                t = t.permute(0, 2, 1)                 # [out, in, k] -> MLX [out, k, in]
        else:
            m = _HF_LAYER.match(k)
            if not m or m.group(3) not in _HF_TO_MLX_SUB:
                raise KeyError(f"no MLX name for {k!r}")
            name = f"{m.group(1)}.blocks.{m.group(2)}.{_HF_TO_MLX_SUB[m.group(3)]}.{m.group(4)}"
        out[name] = t.contiguous()
    return out


def load_weights_dir(path: str, arch: WhisperArch, prefix: str = "") -> Dict[str, torch.Tensor]:
    """model.safetensors first, then model.npz / weights.npz, as ref:scripts/evaluate_model.py:41-56.
    ``prefix`` keeps only the keys that start with it (the reference overlays ``k.startswith('decoder.')`` only,
    ref:scripts/evaluate_model.py:58) BEFORE any name is mapped."""
    def keep(d):
        return {k: v for k, v in d.items() if k.startswith(prefix) or (prefix and k.startswith("model." + prefix))}
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return to_hf_state_dict(keep(load_file(st)), arch)
    for name in ("model.npz", "weights.npz"):
        p = os.path.join(path, name)
        if os.path.exists(p):
            return to_hf_state_dict(keep(dict(np.load(p))), arch)
    raise FileNotFoundError(f"no model.safetensors / model.npz under {path!r} (the reference exits here, "
                            "ref:scripts/transcribe_single.py:36-37)")
