"""The reference's script-level functions with their signatures kept (ref:scripts/evaluate_model.py:20-268,
ref:scripts/transcribe_single.py:10-56), running on libwipa instead of mlx_whisper / editdistance."""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, Optional

import numpy as np

from .archs import arch_from_name
from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, decode
from .metrics import evaluate_batch, phone_error_rate
from .model import WhisperIPA, load_model


def load_checkpoint_model(checkpoint_path: str, base_model: str = "mlx-community/whisper-small-mlx", dtype="float32",
                          base_state_dict=None, max_batch: int = 16) -> WhisperIPA:
    """Base weights + overlay of every ``decoder.*`` tensor of the checkpoint (ref:scripts/evaluate_model.py:20-79)."""
    from .checkpoint import load_weights_dir, to_hf_state_dict
    model = WhisperIPA(arch_from_name(base_model), dtype=dtype, max_batch=max_batch)
    if base_state_dict is not None:
        model.load_state_dict(base_state_dict)
    elif os.path.isdir(base_model):
        model.load_state_dict(load_weights_dir(base_model, model.arch))
    else:
        raise FileNotFoundError(f"base model weights for {base_model!r} are not available offline; pass base_state_dict "
                                "or a local directory")
    try:
        overlay = load_weights_dir(checkpoint_path, model.arch)
    except FileNotFoundError as e:
        print(f"Error: Could not find weights in {checkpoint_path}")       # ref:scripts/transcribe_single.py:36-37
        raise SystemExit(1) from e
    dec = {k: v for k, v in overlay.items() if k.startswith("model.decoder.") or k == "proj_out.weight"}
    model.load_state_dict(dec)
    return model


def transcribe_file(model: WhisperIPA, audio_path: str) -> str:
    """ref:scripts/transcribe_single.py:41-56."""
    audio = pad_or_trim(load_audio(audio_path))
    mel = log_mel_spectrogram(audio, n_mels=model.arch.n_mels)
    mel = mel[None].float()
    audio_features = model.encoder(mel)
    options = DecodingOptions(language="en", without_timestamps=True)
    result = decode(model, audio_features, options)
    return result[0].text.strip()


def evaluate_model(model_path, test_data_path: str, num_samples: Optional[int] = None, model_name: str = "Model",
                   is_checkpoint: bool = False, n_mels: int = 80, base_model: str = "mlx-community/whisper-small-mlx",
                   model: Optional[WhisperIPA] = None) -> Optional[Dict]:
    """ref:scripts/evaluate_model.py:127-232: per-sample transcription of a JSON test set, then evaluate_batch.
    ``model`` lets a caller hand in an already-built model (random-init tests); failures become "" like upstream."""
    with open(test_data_path, "r", encoding="utf-8") as f:
        test_data = json.load(f)
    if num_samples:
        test_data = test_data[:num_samples]
    if model is None:
        model = load_checkpoint_model(model_path, base_model) if is_checkpoint else load_model(model_path)
    if n_mels != model.arch.n_mels:
        raise ValueError(f"--n-mels {n_mels} does not match {model.arch.name} (n_mels={model.arch.n_mels})")
    options = DecodingOptions(language="en", without_timestamps=True)
    references, hypotheses = [], []
    for sample in test_data:
        reference_ipa = sample["ipa_transcription"]
        try:
            audio = pad_or_trim(load_audio(sample["audio_path"]))
            mel = log_mel_spectrogram(audio, n_mels=n_mels)[None].float()
            audio_features = model.encoder(mel)
            hypothesis = decode(model, audio_features, options)[0].text.strip()
        except Exception as e:                                            # ref:scripts/evaluate_model.py:202-204
            print(f"Error transcribing {sample.get('audio_path')}: {e}")
            hypothesis = ""
        references.append(reference_ipa)
        hypotheses.append(hypothesis)
    results = evaluate_batch(references, hypotheses)
    print(f"\n{model_name} Results:\n  PER:  {results['per']:.2f}% ± {results['per_std']:.2f}%")
    return results
