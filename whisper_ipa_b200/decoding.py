"""``DecodingOptions`` / ``DecodingResult`` / ``decode`` with the reference's call shapes
(ref:scripts/evaluate_model.py:168-173,200-201, ref:scripts/transcribe_single.py:49-56,
ref:scripts/train_whisper_ipa.py:338-362).  Greedy, temperature 0, one 30 s window."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Union

import torch

MAX_TARGET = 448

# optional id -> text hook; the Whisper BPE vocabulary is not shipped in this image
_detokenizer: Optional[Callable[[Sequence[int]], str]] = None


def set_detokenizer(fn: Optional[Callable[[Sequence[int]], str]]) -> None:
    global _detokenizer
    _detokenizer = fn


@dataclass(frozen=True)
class DecodingOptions:
    task: str = "transcribe"
    language: Optional[str] = "en"
    temperature: float = 0.0
    sample_len: Optional[int] = None          # default n_text_ctx // 2 = 224 tokens including the prompt
    beam_size: Optional[int] = None
    length_penalty: Optional[float] = None
    without_timestamps: bool = True
    fp16: bool = False
    suppress_tokens: Optional[Sequence[int]] = None
    suppress_blank: bool = True


@dataclass
class DecodingResult:
    audio_features: Optional[torch.Tensor]
    language: str
    tokens: List[int] = field(default_factory=list)
    text: str = ""


def _text(tokens: Sequence[int]) -> str:
    if _detokenizer is not None:
        return _detokenizer(tokens)
    return " ".join(str(t) for t in tokens)          # ids as text: PER over token ids stays well defined


def decode(model, mel_or_features, options: Optional[DecodingOptions] = None) -> Union[DecodingResult, List[DecodingResult]]:
    """mel [B,3000,n_mels] / [3000,n_mels] or encoder output [B,1500,d] -> DecodingResult(s)."""
    o = options or DecodingOptions()
    if o.temperature != 0.0:
        raise NotImplementedError("temperature sampling is outside the reference's evaluation path")
    if o.language not in (None, "en"):
        raise NotImplementedError("the reference decodes every language with the <|en|> prompt")
    x = torch.as_tensor(mel_or_features)
    single = x.dim() == 2
    if single:
        x = x[None]
    if x.shape[-2:] == (1500, model.arch.d_model):
        model.set_audio_features(x)
        feats = x
    else:
        feats = model.encoder(x)
    prompt = model.arch.prompt("en", o.task, o.without_timestamps)
    sample_len = o.sample_len if o.sample_len is not None else MAX_TARGET // 2
    max_new = sample_len - len(prompt)
    begin = list(model.begin_suppress_tokens) if o.suppress_blank else [model.arch.eot]
    ids, lens = model.decode_tokens(prompt, max_new, num_beams=o.beam_size or 1,
                                    length_penalty=1.0 if o.length_penalty is None else o.length_penalty,
                                    suppress=o.suppress_tokens, begin_suppress=begin)
    ids_h, lens_h = ids.cpu().tolist(), lens.cpu().tolist()
    results = []
    for b in range(len(ids_h)):
        toks = ids_h[b][:lens_h[b]]
        results.append(DecodingResult(audio_features=feats[b] if feats is not None else None, language="en",
                                      tokens=toks, text=_text(toks)))
    return results[0] if single else results
