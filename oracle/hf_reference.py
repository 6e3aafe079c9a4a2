"""The parity oracle BASELINE.json names: HF transformers' Whisper, random-init, run on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/whisper_oracle.py header).  HF `transformers` is a library in
the image (present here and on the GPU box); it is *not* part of /root/reference.  This module
builds the seeded random-init model exactly as SURVEY.md §8d specifies and drives the stock
``WhisperFeatureExtractor`` / ``WhisperForConditionalGeneration.generate`` code path, so the
restatement in ``whisper_oracle.py`` and the CUDA path can both be compared against it.
"""
from __future__ import annotations

import logging
import warnings
from typing import Dict, Sequence

import numpy as np
import torch

from .whisper_oracle import ARCHS, BEGIN_SUPPRESS, EOT, prompt_for


def build_hf_model(arch: str, seed: int = 0, init_gain: float = 1.0):
    """torch.manual_seed(seed) + WhisperForConditionalGeneration(WhisperConfig(**arch, ...)).

    ``init_gain`` > 1 multiplies every >=2-D weight except the (fixed, sinusoidal) encoder positions
    after the default init; used by the sensitivity tests, because a std-0.02 random model barely reacts to its audio
    input and would let an encoder bug slip through an ids-only comparison.
    """
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    logging.getLogger("transformers").setLevel(logging.ERROR)
    cfg = WhisperConfig(**ARCHS[arch], decoder_start_token_id=50258, eos_token_id=EOT, pad_token_id=EOT,
                        bos_token_id=EOT, begin_suppress_tokens=list(BEGIN_SUPPRESS))
    torch.manual_seed(seed)
    model = WhisperForConditionalGeneration(cfg).eval()
    if init_gain != 1.0:
        with torch.no_grad():
            for n, p in model.named_parameters():
                if p.dim() >= 2 and "embed_positions" not in n:
                    p.mul_(init_gain)
    return model


def state_dict_f32(model) -> Dict[str, torch.Tensor]:
    sd = {k: v.detach().to(torch.float32).contiguous() for k, v in model.state_dict().items()}
    if "proj_out.weight" not in sd:
        sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    return sd


def hf_log_mel(audio: np.ndarray, n_mels: int) -> torch.Tensor:
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=n_mels)
    return fe(list(audio), sampling_rate=16000, return_tensors="pt").input_features


def hf_generate(model, feats: torch.Tensor, arch: str, max_new: int, num_beams: int = 1,
                prompt: Sequence[int] | None = None) -> torch.Tensor:
    """Greedy / beam ids through the stock generate(); prompt passed as decoder_input_ids
    (HF:models/whisper/generation_whisper.py:1676-1682).  Returns int64[B, <=max_new], prompt and EOS
    stripped, right-padded with pad_token_id (= EOT)."""
    prompt = list(prompt) if prompt is not None else prompt_for(arch)
    B = feats.shape[0]
    ids = torch.tensor([prompt] * B, dtype=torch.long)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(decoder_input_ids=ids, max_new_tokens=max_new, do_sample=False)
        if num_beams > 1:
            kw.update(num_beams=num_beams, length_penalty=1.0, early_stopping=False)
        out = model.generate(feats, **kw)
    return out


def hf_gpu_fp32_reference(model, audio: np.ndarray, prompt: Sequence[int], max_new: int, logit_steps: int,
                          device: str = "cuda", logit_rows: int | None = None):
    """The fp32 oracle on the GPU (HF in fp32 with TF32 disabled): stock feature extractor on the host,
    then the stock encoder, ``generate`` (greedy) and one teacher-forced decoder pass over the prompt + the first
    ``logit_steps - 1`` greedy tokens, all in true fp32 (cuBLAS fp32 kernels, no TF32, SDPA math).

    Returns a dict of CPU tensors: ``mel`` [B, n_mels, 3000], ``enc`` [B, 1500, d], ``ids`` int64 [B, <= max_new] (prompt and
    EOS stripped, EOT-padded as ``generate`` returns them), ``tokens`` int64 [B, P + logit_steps - 1] (the teacher-forced
    input) and ``logits`` fp32 [rows, P + logit_steps - 1, V] for the first ``logit_rows`` utterances (all when None).
    """
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        mel = hf_log_mel(audio, model.config.num_mel_bins)
        B = mel.shape[0]
        m = model.to(device).float().eval()
        p = torch.tensor([list(prompt)] * B, dtype=torch.long, device=device)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            feats = mel.to(device)
            enc = torch.cat([m.model.encoder(feats[i:i + 32]).last_hidden_state for i in range(0, B, 32)])
            ids = m.generate(feats, decoder_input_ids=p, max_new_tokens=max_new, do_sample=False)
            n_extra = max(0, min(logit_steps - 1, ids.shape[1]))
            toks = torch.cat([p, ids[:, :n_extra]], dim=1)
            rows = B if logit_rows is None else min(B, logit_rows)
            logits = torch.cat([m(encoder_outputs=(enc[i:i + 16],), decoder_input_ids=toks[i:i + 16]).logits.float()
                                for i in range(0, rows, 16)])
        out = {"mel": mel, "enc": enc.cpu(), "ids": ids.cpu(), "tokens": toks.cpu(), "logits": logits.cpu()}
    finally:
        model.to("cpu")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out
