"""Whisper architectures the reference scripts select by base-model name (ref:scripts/evaluate_model.py:281-286,
ref:scripts/transcribe_single.py:12, ref:scripts/train_whisper_ipa.py:517) and the special-token ids its prompt uses
(ref:scripts/ipa_data_loader.py:106-120, ref:WHISPER_IPA_RESEARCH_STANDALONE.md:315-338)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List


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

    def prompt(self, language: str = "en", task: str = "transcribe", without_timestamps: bool = True) -> List[int]:
        """<|sot|><|lang|><|task|>[<|notimestamps|>] — the reference always decodes with language="en"."""
        if language != "en":
            raise ValueError("only the reference's language='en' prompt is wired (ref:scripts/evaluate_model.py:171)")
        shift = 1 if self.is_v3 else 0
        task_id = {"transcribe": 50359, "translate": 50358}[task] + shift
        out = [self.sot, 50259, task_id]
        if without_timestamps:
            out.append(50363 + shift)
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
    s = base_model.lower()
    for key in ("large-v3", "medium", "small", "base", "tiny"):
        if key in s:
            return ARCHS[key]
    raise ValueError(f"cannot infer a Whisper architecture from {base_model!r}")
