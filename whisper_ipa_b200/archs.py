"""Whisper architectures the reference scripts select by base-model name (ref:scripts/evaluate_model.py:281-286,
ref:scripts/transcribe_single.py:12, ref:scripts/train_whisper_ipa.py:517) and the special-token ids its prompt uses
(ref:scripts/ipa_data_loader.py:106-120, ref:WHISPER_IPA_RESEARCH_STANDALONE.md:315-338)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List


# Whisper's language codes in token order: <|en|> = 50259, <|zh|> = 50260, ... (99 languages; large-v3 appends "yue").
# Same order as HF:models/whisper/tokenization_whisper.py LANGUAGES (checked in tests/test_host_logic.py).
LANGUAGES = (
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue",
)
FIRST_LANGUAGE_TOKEN = 50259

# The multilingual tokenizer's "non-speech" symbols (brackets, quotes, musical notes, ...): the BPE ids below <|endoftext|>
# that Whisper masks at every step when suppress_tokens == "-1" (the default of the reference's
# DecodingOptions(language="en", without_timestamps=True), ref:scripts/evaluate_model.py:170-173).  Same ids for every
# multilingual vocabulary (tiny ... large-v3); equal to the entries < 50257 of HF:models/whisper/configuration_whisper.py
# NON_SPEECH_TOKENS_MULTI (checked in tests/test_host_logic.py).
NON_SPEECH_TOKENS = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873, 893, 902,
    918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647, 7273, 9061, 9383,
    10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362, 18956, 20075, 21675,
    22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870, 50254,
)


@dataclass(frozen=True)
class WhisperArch:
    name: str
    d_model: int
    enc_layers: int
    dec_layers: int
    heads: int
    ffn: int
    n_mels: int
    vocab: int

    @property
    def is_v3(self) -> bool:
        return self.vocab == 51866

    # special tokens (multilingual vocabularies)
    @property
    def eot(self) -> int:
        return 50257

    @property
    def sot(self) -> int:
        return 50258

    @property
    def n_languages(self) -> int:
        return 100 if self.is_v3 else 99

    def language_token(self, language: str) -> int:
        """<|xx|> id of a language code (50259 + its position in LANGUAGES)."""
        try:
            i = LANGUAGES.index(language)
        except ValueError:
            raise ValueError(f"unknown language code {language!r}") from None
        if i >= self.n_languages:
            raise ValueError(f"language {language!r} is not in this model's vocabulary")
        return FIRST_LANGUAGE_TOKEN + i

    def language_of_token(self, token: int) -> str:
        i = int(token) - FIRST_LANGUAGE_TOKEN
        if not 0 <= i < self.n_languages:
            raise ValueError(f"token {token} is not a language token")
        return LANGUAGES[i]

    # <|translate|>, <|transcribe|>, <|startoflm|>, <|startofprev|>, <|nospeech|>, <|notimestamps|>: large-v3 inserted one more
    # language token in front of them, so they sit one id higher there
    @property
    def _shift(self) -> int:
        return 1 if self.is_v3 else 0

    @property
    def translate(self) -> int:
        return 50358 + self._shift

    @property
    def transcribe(self) -> int:
        return 50359 + self._shift

    @property
    def sot_lm(self) -> int:
        return 50360 + self._shift

    @property
    def sot_prev(self) -> int:
        return 50361 + self._shift

    @property
    def no_speech(self) -> int:
        return 50362 + self._shift

    @property
    def no_timestamps(self) -> int:
        return 50363 + self._shift

    @property
    def timestamp_begin(self) -> int:
        return 50364 + self._shift

    def default_suppress_tokens(self) -> List[int]:
        """What ``suppress_tokens="-1"`` expands to (mlx_whisper / openai-whisper ``_get_suppress_tokens``): the non-speech
        symbols plus <|transcribe|>, <|translate|>, <|startoftranscript|>, <|startofprev|>, <|startoflm|> and <|nospeech|>."""
        return sorted(set(NON_SPEECH_TOKENS) | {self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm,
                                                self.no_speech})

    def resolve_suppress_tokens(self, spec) -> List[int]:
        """``DecodingOptions.suppress_tokens``: None / "" -> nothing; "-1" or a list containing -1 -> the default set (plus
        the list's other ids); a comma-separated string or an iterable of ids -> those ids."""
        if spec is None:
            return []
        if isinstance(spec, str):
            spec = [int(t) for t in spec.split(",") if t.strip()]
        ids = [int(t) for t in spec]
        out = set(t for t in ids if t >= 0)
        if any(t == -1 for t in ids):
            out |= set(self.default_suppress_tokens())
        bad = [t for t in out if t >= self.vocab]
        if bad:
            raise ValueError(f"suppress_tokens outside the vocabulary: {bad[:5]}")
        return sorted(out)

    def prompt(self, language: str = "en", task: str = "transcribe", without_timestamps: bool = True) -> List[int]:
        """<|sot|><|lang|><|task|>[<|notimestamps|>] — the evaluation scripts always decode with language="en"
        (ref:scripts/evaluate_model.py:171); training-time validation detects the language (ref:scripts/train_whisper_ipa.py:339)."""
        task_id = {"transcribe": self.transcribe, "translate": self.translate}[task]
        out = [self.sot, self.language_token(language), task_id]
        if without_timestamps:
            out.append(self.no_timestamps)
        return out

    def hf_config_kwargs(self) -> Dict[str, int]:
        return dict(d_model=self.d_model, encoder_layers=self.enc_layers, decoder_layers=self.dec_layers,
                    encoder_attention_heads=self.heads, decoder_attention_heads=self.heads,
                    encoder_ffn_dim=self.ffn, decoder_ffn_dim=self.ffn, num_mel_bins=self.n_mels,
                    vocab_size=self.vocab)


ARCHS: Dict[str, WhisperArch] = {a.name: a for a in (
    WhisperArch("tiny", 384, 4, 4, 6, 1536, 80, 51865),
    WhisperArch("base", 512, 6, 6, 8, 2048, 80, 51865),
    WhisperArch("small", 768, 12, 12, 12, 3072, 80, 51865),
    WhisperArch("medium", 1024, 24, 24, 16, 4096, 80, 51865),
    WhisperArch("large-v3", 1280, 32, 32, 20, 5120, 128, 51866),
)}


def arch_from_name(base_model: str) -> WhisperArch:
    """Map a reference-style model id ("mlx-community/whisper-small-mlx", "openai/whisper-large-v3", "small") to an arch."""
    import os
    # the last path component first: a local directory such as /data/base/whisper-small-mlx names "small", not "base"
    for s in (os.path.basename(base_model.rstrip("/")).lower(), base_model.lower()):
        for key in ("large-v3", "medium", "small", "base", "tiny"):
            if key in s:
                return ARCHS[key]
    raise ValueError(f"cannot infer a Whisper architecture from {base_model!r}")
