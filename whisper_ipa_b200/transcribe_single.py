"""``scripts/transcribe_single.py`` of the reference on libwipa (ref:scripts/transcribe_single.py:10-68): large-v3 base model,
``decoder.*`` overlay from ``model.safetensors`` (exit 1 when it is missing), one file -> IPA string.

    python -m whisper_ipa_b200.transcribe_single [CHECKPOINT_DIR [AUDIO_FILE]]
"""
from __future__ import annotations

import os
import sys

from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, decode
from .evaluate_model import _try_load_detokenizer, load_base_model
from .model import WhisperIPA

DEFAULT_BASE = "mlx-community/whisper-large-v3-mlx"


def load_checkpoint_model(checkpoint_path: str, base_model: str = DEFAULT_BASE, dtype="float32", max_batch: int = 1,
                          base_state_dict=None) -> WhisperIPA:
    """ref:scripts/transcribe_single.py:10-39 - unlike evaluate_model's loader, a checkpoint without ``model.safetensors``
    is an ERROR and the process exits with status 1 (:36-37)."""
    from .checkpoint import load_weights_dir
    print(f"Loading base model architecture: {base_model}")
    model = load_base_model(base_model, dtype=dtype, max_batch=max_batch, base_state_dict=base_state_dict)
    weights_path = os.path.join(checkpoint_path, "model.safetensors")
    if not os.path.exists(weights_path):
        print(f"ERROR: No weights found at {weights_path}")
        sys.exit(1)
    print(f"Loading trained weights from: {weights_path}")
    overlay = load_weights_dir(checkpoint_path, model.arch, prefix="decoder.")
    print(f"Found {len(overlay)} decoder parameters to load")
    model.load_state_dict(overlay)
    _try_load_detokenizer(checkpoint_path)
    print("✓ Decoder weights loaded successfully")
    return model


def transcribe_file(model: WhisperIPA, audio_path: str) -> str:
    """ref:scripts/transcribe_single.py:41-56 (the reference hard-codes n_mels=128 for large-v3; here it follows the model)."""
    from .decoding import require_detokenizer
    require_detokenizer()
    print(f"Transcribing {audio_path}...")
    audio = load_audio(audio_path)
    audio = pad_or_trim(audio)
    mel = log_mel_spectrogram(audio, n_mels=model.arch.n_mels)
    mel = mel[None].float()
    decode_options = DecodingOptions(language="en", without_timestamps=True)   # IPA is treated as English for the tokenizer
    audio_features = model.encoder(mel)
    result = decode(model, audio_features, decode_options)
    return result[0].text.strip()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    checkpoint = argv[0] if len(argv) > 0 else "checkpoints/whisper-ipa/checkpoint-8000"
    audio_file = argv[1] if len(argv) > 1 else "4.wav"
    model = load_checkpoint_model(checkpoint)
    transcription = transcribe_file(model, audio_file)
    print("\n" + "=" * 50)
    print(f"Audio: {audio_file}")
    print(f"Prediction: {transcription}")
    print("=" * 50)
    return transcription


if __name__ == "__main__":
    main()
