"""Audio front end behind the reference's call sites: load_audio / pad_or_trim / log_mel_spectrogram
(ref:scripts/evaluate_model.py:187-189, ref:scripts/transcribe_single.py:43-45, ref:scripts/ipa_data_loader.py:48,80-84).

The log-mel arithmetic runs in libwipa's CUDA kernels (csrc/logmel.cu); this module only moves buffers.
"""
from __future__ import annotations

import wave
from typing import Sequence, Union

import numpy as np
import torch

from . import _lib

SAMPLE_RATE = 16000
N_SAMPLES = 480000
N_FRAMES = 3000


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def load_audio(path: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """PCM WAV -> mono float32 at 16 kHz.  The reference shells out to ffmpeg (absent here); other containers raise."""
    with wave.open(path, "rb") as w:
        n_ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    if n_ch > 1:
        x = x.reshape(-1, n_ch).mean(axis=1)
    if rate != sr:
        from scipy.signal import resample_poly
        g = np.gcd(rate, sr)
        x = resample_poly(x, sr // g, rate // g).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def pad_or_trim(array, length: int = N_SAMPLES, axis: int = -1):
    """Zero-pad or cut to 30 s (ref:scripts/evaluate_model.py:188)."""
    if isinstance(array, torch.Tensor):
        n = array.shape[axis]
        if n > length:
            array = array.narrow(axis, 0, length)
        elif n < length:
            pad = [0, 0] * array.dim()
            pad[2 * (array.dim() - 1 - (axis % array.dim())) + 1] = length - n
            array = torch.nn.functional.pad(array, pad)
        return array
    array = np.asarray(array)
    n = array.shape[axis]
    if n > length:
        array = array.take(indices=range(length), axis=axis)
    elif n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        array = np.pad(array, widths)
    return array


def _to_device_audio(audio) -> torch.Tensor:
    if isinstance(audio, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    elif isinstance(audio, torch.Tensor):
        t = audio.to(torch.float32)
    else:
        t = torch.from_numpy(np.stack([np.asarray(a, dtype=np.float32) for a in audio]))
    if t.dim() == 1:
        t = t[None]
    if t.dim() != 2 or t.shape[1] != N_SAMPLES:
        raise ValueError(f"expected audio of shape [B, {N_SAMPLES}] (use pad_or_trim), got {tuple(t.shape)}")
    if not t.is_cuda:
        t = t.pin_memory().cuda(non_blocking=True) if torch.cuda.is_available() else t.cuda()
    return t.contiguous()


def log_mel_features(audio, n_mels: int = 80) -> torch.Tensor:
    """audio [B, 480000] (numpy / torch, host or device) -> device f32 [B, n_mels, 3000] (HF layout)."""
    a = _to_device_audio(audio)
    mel = torch.empty((a.shape[0], n_mels, N_FRAMES), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().wipa_logmel(a.data_ptr(), a.shape[0], n_mels, mel.data_ptr(), _stream()), "wipa_logmel")
    return mel


def log_mel_spectrogram(audio, n_mels: int = 80) -> torch.Tensor:
    """Reference-shaped result: [3000, n_mels] for one clip, [B, 3000, n_mels] for a batch
    (ref:scripts/ipa_data_loader.py:82-84)."""
    single = (np.ndim(audio) == 1) if not isinstance(audio, torch.Tensor) else audio.dim() == 1
    mel = log_mel_features(audio, n_mels).transpose(1, 2)
    return mel[0] if single else mel


class FeatureExtractor:
    """HF-face: ``fe(list_of_clips, sampling_rate=16000, return_tensors="pt").input_features`` -> f32 [B, n_mels, 3000]
    (HF:models/whisper/feature_extraction_whisper.py:189-342); clips are zero-padded / truncated to 30 s like HF."""

    def __init__(self, feature_size: int = 80):
        self.feature_size = feature_size

    def __call__(self, raw_speech: Union[np.ndarray, Sequence[np.ndarray]], sampling_rate: int = SAMPLE_RATE,
                 return_tensors: str = "pt"):
        if sampling_rate != SAMPLE_RATE:
            raise ValueError(f"sampling_rate must be {SAMPLE_RATE}")
        if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 1:
            raw_speech = [raw_speech]
        batch = np.stack([pad_or_trim(np.asarray(x, dtype=np.float32)) for x in raw_speech])
        feats = log_mel_features(batch, self.feature_size)

        class _Out(dict):
            __getattr__ = dict.__getitem__
        return _Out(input_features=feats)
