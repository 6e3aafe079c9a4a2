"""``DecodingOptions`` / ``DecodingResult`` / ``decode`` with the reference's call shapes
(ref:scripts/evaluate_model.py:168-173,200-201, ref:scripts/transcribe_single.py:49-56,
ref:scripts/train_whisper_ipa.py:338-362).  Greedy, temperature 0, one 30 s window."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch

from .archs import FIRST_LANGUAGE_TOKEN, LANGUAGES

MAX_TARGET = 448

# id -> text.  The Whisper BPE vocabulary is not shipped in this image, so it comes from the
# checkpoint directory (load_detokenizer) or from the caller (set_detokenizer).  Without one, asking for `.text` RAISES:
# ids printed as text would turn every string PER into a silent ~100 %.  The id-level pipeline (synthetic configs,
# bench.py) opts in to ids-as-text explicitly with set_detokenizer("ids").
_detokenizer: Union[None, str, Callable[[Sequence[int]], str]] = None


def set_detokenizer(fn: Union[None, str, Callable[[Sequence[int]], str]]) -> None:
    """fn(ids) -> str; the string "ids" selects space-joined ids (id-level scoring only); None clears it."""
    global _detokenizer
    if isinstance(fn, str) and fn != "ids":
        raise ValueError('set_detokenizer takes a callable, "ids" or None')
    _detokenizer = fn


def load_detokenizer(path: str, eot: int = 50257) -> Callable[[Sequence[int]], str]:
    """Build (and register) the id -> text function from tokenizer files next to a checkpoint:
      * ``multilingual.tiktoken`` / ``gpt2.tiktoken`` (the asset mlx_whisper / openai-whisper ship: one
        ``base64(token bytes) rank`` pair per line) - decoded here: bytes of the ids below <|endoftext|> joined, UTF-8 with
        replacement, exactly what tiktoken's ``decode`` does once the special tokens are filtered out
        (mlx_whisper.tokenizer.Tokenizer.decode drops ids >= timestamp_begin; the reference reads ``result.text``, which
        is decoded from the tokens before the first EOT, ref:scripts/evaluate_model.py:201);
      * an HF tokenizer directory (``tokenizer.json`` or ``vocab.json`` + ``merges.txt``) through
        ``transformers.WhisperTokenizerFast/WhisperTokenizer`` with ``skip_special_tokens=True``."""
    import base64
    import os
    fn: Optional[Callable[[Sequence[int]], str]] = None
    files = [path] if os.path.isfile(path) else [os.path.join(path, n) for n in ("multilingual.tiktoken", "gpt2.tiktoken")]
    for f in files:
        if os.path.isfile(f) and f.endswith(".tiktoken"):
            table = {}
            with open(f, "rb") as fh:
                for line in fh:
                    if line.strip():
                        tok, rank = line.split()
                        table[int(rank)] = base64.b64decode(tok)

            def fn(ids, _t=table, _eot=eot):
                return b"".join(_t[int(i)] for i in ids if int(i) < _eot and int(i) in _t).decode("utf-8", errors="replace")
            break
    if fn is None and os.path.isdir(path) and any(os.path.exists(os.path.join(path, n)) for n in ("tokenizer.json", "vocab.json")):
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(path, local_files_only=True)

        def fn(ids, _tok=tok):
            return _tok.decode([int(i) for i in ids], skip_special_tokens=True)
    if fn is None:
        raise FileNotFoundError(f"no multilingual.tiktoken / tokenizer.json / vocab.json under {path!r}")
    set_detokenizer(fn)
    return fn


@dataclass(frozen=True)
class DecodingOptions:
    task: str = "transcribe"
    language: Optional[str] = "en"
    temperature: float = 0.0
    sample_len: Optional[int] = None          # tokens to sample AFTER the prompt; default n_text_ctx // 2 = 224 (mlx_whisper
                                              # runs `for i in range(sample_len)`), capped by the 448 target positions
    beam_size: Optional[int] = None
    length_penalty: Optional[float] = None
    without_timestamps: bool = True
    fp16: bool = False
    suppress_tokens: Union[None, str, Sequence[int]] = "-1"      # "-1": non-speech symbols + special tokens (mlx_whisper default)
    suppress_blank: bool = True


class DecodingResult:
    """mlx_whisper.decoding.DecodingResult's fields the reference reads (`.text`, ref:scripts/evaluate_model.py:201) plus
    `.tokens` / `.language` / `.audio_features`.  `.text` is detokenized on first access, so id-level callers never need
    a vocabulary - and string callers get an exception, not ids, when none is registered."""
    __slots__ = ("audio_features", "language", "tokens", "_text")

    def __init__(self, audio_features: Optional[torch.Tensor], language: str, tokens: Optional[List[int]] = None,
                 text: Optional[str] = None):
        self.audio_features, self.language = audio_features, language
        self.tokens = list(tokens) if tokens is not None else []
        self._text = text

    @property
    def text(self) -> str:
        if self._text is None:
            self._text = _text(self.tokens)
        return self._text

    @text.setter
    def text(self, value: str) -> None:
        self._text = value

    def __repr__(self) -> str:
        return f"DecodingResult(language={self.language!r}, tokens={self.tokens!r})"


def require_detokenizer() -> None:
    """Raise unless text output is possible.  Called by the string-level entry points BEFORE any work, so the failure is
    never swallowed by the reference's per-sample ``except`` (which would turn it into empty hypotheses and a PER of 100 %)."""
    if _detokenizer is None:
        raise RuntimeError("no detokenizer is registered: `.text` needs the Whisper vocabulary (decoding.load_detokenizer(<dir "
                           "with multilingual.tiktoken or tokenizer.json>) or set_detokenizer(fn)); for id-level scoring "
                           'opt in with set_detokenizer("ids") or read `.tokens`')


def _text(tokens: Sequence[int]) -> str:
    require_detokenizer()
    if _detokenizer == "ids":
        return " ".join(str(t) for t in tokens)      # ids as text: PER over token ids stays well defined
    return _detokenizer(tokens)


def _encode(model, mel_or_features):
    x = torch.as_tensor(mel_or_features)
    single = x.dim() == 2
    if single:
        x = x[None]
    if x.shape[-2:] == (1500, model.arch.d_model):
        model.set_audio_features(x)
        feats = x
    else:
        feats = model.encoder(x)
    return feats, single


def _detect_cached(model, n: int) -> Tuple[List[str], List[Dict[str, float]]]:
    """Language of each of the n utterances whose encoder output the model holds: ONE decoder step on <|sot|>, every
    logit outside the language tokens masked out, argmax and softmax over the rest (mlx_whisper.decoding.detect_language,
    HF:models/whisper/generation_whisper.py:1610 detect_language do the same)."""
    arch = model.arch
    logits = model.teacher_forced_logits(torch.full((n, 1), arch.sot, dtype=torch.int64))[:, 0].float()
    lo, hi = FIRST_LANGUAGE_TOKEN, FIRST_LANGUAGE_TOKEN + arch.n_languages
    lang_logits = logits[:, lo:hi]
    best = lang_logits.argmax(dim=-1).cpu().tolist()
    probs = torch.softmax(lang_logits, dim=-1).cpu()
    langs = [LANGUAGES[i] for i in best]
    return langs, [{LANGUAGES[j]: float(probs[b, j]) for j in range(arch.n_languages)} for b in range(n)]


def detect_language(model, mel_or_features) -> Tuple[Union[str, List[str]], Union[Dict[str, float], List[Dict[str, float]]]]:
    """mel [B,3000,n_mels] / [3000,n_mels] or encoder output [B,1500,d] -> (language code(s), per-language probabilities).
    What ``DecodingOptions(language=None)`` runs first (ref:scripts/train_whisper_ipa.py:338-343)."""
    feats, single = _encode(model, mel_or_features)
    langs, probs = _detect_cached(model, feats.shape[0])
    return (langs[0], probs[0]) if single else (langs, probs)


def decode(model, mel_or_features, options: Optional[DecodingOptions] = None) -> Union[DecodingResult, List[DecodingResult]]:
    """mel [B,3000,n_mels] / [3000,n_mels] or encoder output [B,1500,d] -> DecodingResult(s).
    ``language=None`` detects the language of every utterance first and decodes each with its own <|lang|> prompt
    (utterances are grouped by language: the C ABI takes one prompt per call)."""
    o = options or DecodingOptions()
    if o.temperature != 0.0:
        raise NotImplementedError("temperature sampling is outside the reference's evaluation path")
    feats, single = _encode(model, mel_or_features)
    B = feats.shape[0]
    sample_len = o.sample_len if o.sample_len is not None else MAX_TARGET // 2
    begin = list(model.begin_suppress_tokens) if o.suppress_blank else [model.arch.eot]
    suppress = model.arch.resolve_suppress_tokens(o.suppress_tokens)

    def run(language: str):
        prompt = model.arch.prompt(language, o.task, o.without_timestamps)
        # mlx_whisper samples `sample_len` tokens after the prompt (`for i in range(sample_len)`); the 448 target positions cap it
        max_new = min(sample_len, MAX_TARGET - len(prompt))
        ids, lens = model.decode_tokens(prompt, max_new, num_beams=o.beam_size or 1,
                                        length_penalty=1.0 if o.length_penalty is None else o.length_penalty,
                                        suppress=suppress, begin_suppress=begin)
        return ids.cpu().tolist(), lens.cpu().tolist()

    langs = [o.language] * B if o.language is not None else _detect_cached(model, B)[0]
    tokens: List[List[int]] = [[] for _ in range(B)]
    groups: Dict[str, List[int]] = {}
    for b, lang in enumerate(langs):
        groups.setdefault(lang, []).append(b)
    if len(groups) == 1:                                   # the usual case: the cached encoder output serves as is
        ids_h, lens_h = run(langs[0])
        for b in range(B):
            tokens[b] = ids_h[b][:lens_h[b]]
    else:
        for lang, rows in groups.items():
            model.set_audio_features(feats[rows])
            ids_h, lens_h = run(lang)
            for i, b in enumerate(rows):
                tokens[b] = ids_h[i][:lens_h[i]]
    results = [DecodingResult(audio_features=feats[b] if feats is not None else None, language=langs[b], tokens=tokens[b])
               for b in range(B)]
    return results[0] if single else results
