"""whisper_ipa_b200 — B200-native (sm_100a) hot path of barathanaslan/whisper-ipa: batched Whisper transcription + PER.

Importing the package does not need a GPU; every compute entry point goes through libwipa.so (CUDA) and raises when the
library or a B200 is missing — there is no CPU fallback."""
from .archs import ARCHS, WhisperArch, arch_from_name
from .audio import FeatureExtractor, load_audio, log_mel_features, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult, decode, detect_language, load_detokenizer, set_detokenizer
from .evaluate_model import (compare_models, evaluate_model, load_base_model, load_checkpoint_model, transcribe_batched,
                             transcribe_with_model)
from .transcribe_single import transcribe_file
from .metrics import (evaluate_batch, normalize_ipa_for_comparison, phone_error_rate, phone_error_rates, tokenize_ipa)
from .model import WhisperIPA, load_model

__all__ = ["detect_language", "ARCHS", "WhisperArch", "arch_from_name", "FeatureExtractor", "load_audio", "log_mel_features",
           "log_mel_spectrogram", "pad_or_trim", "DecodingOptions", "DecodingResult", "decode", "evaluate_batch",
           "normalize_ipa_for_comparison", "phone_error_rate", "phone_error_rates", "tokenize_ipa", "WhisperIPA",
           "load_model", "load_detokenizer", "set_detokenizer", "compare_models", "evaluate_model", "load_base_model",
           "load_checkpoint_model", "transcribe_batched", "transcribe_with_model", "transcribe_file"]
