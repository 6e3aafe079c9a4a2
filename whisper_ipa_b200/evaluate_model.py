"""``scripts/evaluate_model.py`` of the reference on libwipa: the same functions, signatures, prints and CLI flags
(ref:scripts/evaluate_model.py:20-346), with the per-sample loop of the checkpoint branch (:179-209) run as micro-batches:
a worker pool reads the audio files while the GPU transcribes the previous batch (ingest.py), log-mel / encoder /
greedy decode run for the whole batch, PER and PFER are scored on the GPU by ``evaluate_batch``.

    python -m whisper_ipa_b200.evaluate_model --checkpoint CKPT --base-model BASE_DIR --test-data test.json --n-mels 80

Base weights: the reference downloads ``mlx-community/whisper-*-mlx`` from the hub; here ``--base-model`` is a local directory
(MLX- or HF-named ``model.safetensors`` / ``weights.npz``), or a hub id resolved under ``$WIPA_MODEL_DIR/<last path part>``.
The tokenizer (``multilingual.tiktoken`` or HF tokenizer files) is looked up in the checkpoint / base directory.
"""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, List, Optional

from . import decoding
from .archs import arch_from_name
from .decoding import DecodingOptions, decode
from .metrics import evaluate_batch, pfer_available, phone_error_rate, phone_feature_error_rate
from .model import WhisperIPA

DEFAULT_BASE = "mlx-community/whisper-small-mlx"


def resolve_model_dir(name: str) -> Optional[str]:
    """A local directory for a model id: the path itself, or $WIPA_MODEL_DIR/<name> / <basename> (there is no hub access)."""
    if os.path.isdir(name):
        return name
    root = os.environ.get("WIPA_MODEL_DIR")
    if root:
        for cand in (os.path.join(root, name), os.path.join(root, os.path.basename(name.rstrip("/")))):
            if os.path.isdir(cand):
                return cand
    return None


def _try_load_detokenizer(*dirs: Optional[str]) -> None:
    if decoding._detokenizer is not None:
        return
    for d in dirs:
        if d and os.path.isdir(d):
            try:
                decoding.load_detokenizer(d)
                return
            except FileNotFoundError:
                continue


def load_base_model(base_model: str, dtype="float32", max_batch: int = 16, base_state_dict=None) -> WhisperIPA:
    """``mlx_whisper.load_models.load_model(base_model)`` + ``model.set_dtype`` (ref:scripts/evaluate_model.py:33-37)."""
    from .checkpoint import load_weights_dir
    model = WhisperIPA(arch_from_name(base_model), dtype=dtype, max_batch=max_batch)
    if base_state_dict is not None:
        model.load_state_dict(base_state_dict)
        return model
    d = resolve_model_dir(base_model)
    if d is None:
        model.close()
        raise FileNotFoundError(f"base model weights for {base_model!r} are not available offline: pass a local directory or set "
                                "WIPA_MODEL_DIR to a folder that holds it")
    model.load_state_dict(load_weights_dir(d, model.arch))
    _try_load_detokenizer(d)
    return model


def load_checkpoint_model(checkpoint_path: str, base_model: str = DEFAULT_BASE, dtype="float32", max_batch: int = 16,
                          base_state_dict=None) -> WhisperIPA:
    """Base weights + overlay of every key that starts with ``decoder.`` from ``model.safetensors`` / ``model.npz`` of the
    checkpoint; no weights there -> WARNING and the base model (ref:scripts/evaluate_model.py:20-79)."""
    from .checkpoint import load_weights_dir
    print(f"Loading base model architecture: {base_model}")
    model = load_base_model(base_model, dtype=dtype, max_batch=max_batch, base_state_dict=base_state_dict)
    try:
        overlay = load_weights_dir(checkpoint_path, model.arch, prefix="decoder.")
    except FileNotFoundError:
        print(f"WARNING: No weights found at {checkpoint_path}, using base model")
        return model
    print(f"Loading trained weights from: {checkpoint_path}")
    print(f"Found {len(overlay)} decoder parameters to load")
    model.load_state_dict(overlay)
    _try_load_detokenizer(checkpoint_path)
    print("✓ Decoder weights loaded successfully")
    return model


_model_cache: Dict[str, WhisperIPA] = {}


def transcribe_with_model(model_path: str, audio_path: str, language: str = "en", is_checkpoint: bool = False,
                          model: Optional[WhisperIPA] = None) -> str:
    """Long-form transcription with the base model (``mlx_whisper.transcribe(audio_path, path_or_hf_repo=model_path,
    language=language, word_timestamps=False)``); any failure prints and returns "" (ref:scripts/evaluate_model.py:82-124)."""
    from .transcribe import transcribe
    try:
        if model is None:
            model = _model_cache.get(model_path)
            if model is None:
                model = load_checkpoint_model(model_path) if is_checkpoint else load_base_model(model_path)
                _model_cache[model_path] = model
        result = transcribe(audio_path, model, language=language, word_timestamps=False)
        return result["text"].strip()
    except Exception as e:
        print(f"Error transcribing {audio_path}: {e}")
        return ""


def transcribe_batched(model: WhisperIPA, audio_paths: List[str], n_mels: int, batch_size: Optional[int] = None,
                       options: Optional[DecodingOptions] = None, ingest=None) -> List[str]:
    """The checkpoint branch's per-sample body (ref:scripts/evaluate_model.py:185-204) for all files, micro-batched:
    load_audio -> pad_or_trim -> log_mel_spectrogram -> model.encoder -> decode(..., language="en", without_timestamps=True)
    -> ``result.text.strip()``.  A file that fails gives "" like the reference's ``except``."""
    from .audio import log_mel_features
    from .ingest import AudioIngest
    decoding.require_detokenizer()
    options = options or DecodingOptions(language="en", without_timestamps=True)
    own = ingest is None
    ingest = ingest or AudioIngest(device=model.device)
    hyps = [""] * len(audio_paths)
    try:
        for idx, audio, errors in ingest.iter_batches(audio_paths, batch_size or model.max_batch):
            try:
                mel = log_mel_features(audio, n_mels)
                audio_features = model.encoder(mel)
                results = decode(model, audio_features, options)
                for j, i in enumerate(idx):
                    if errors[j] is not None:
                        print(f"\nError transcribing {audio_paths[i]}: {errors[j]}")
                    else:
                        hyps[i] = results[j].text.strip()
            except Exception as e:
                for i in idx:
                    print(f"\nError transcribing {audio_paths[i]}: {e}")
    finally:
        if own:
            ingest.close()
    return hyps


def evaluate_model(model_path: str, test_data_path: str, num_samples: int = None, model_name: str = "Model",
                   is_checkpoint: bool = False, n_mels: int = 80, base_model: str = DEFAULT_BASE,
                   model: Optional[WhisperIPA] = None, batch_size: Optional[int] = None, dtype="float32"):
    """ref:scripts/evaluate_model.py:127-232.  ``model`` hands in an already-built model (tests with random-init weights);
    ``batch_size`` / ``dtype`` size and type the GPU context when one is built here."""
    print("=" * 70)
    print(f"Evaluating {model_name}")
    print("=" * 70)
    print(f"\nLoading test data: {test_data_path}")
    with open(test_data_path) as f:
        test_data = json.load(f)
    if num_samples:
        test_data = test_data[:num_samples]
        print(f"Evaluating on {num_samples} samples")
    else:
        print(f"Evaluating on all {len(test_data)} samples")
    print(f"\nModel: {model_path}")

    references = [sample["ipa_transcription"] for sample in test_data]
    paths = [sample["audio_path"] for sample in test_data]
    if is_checkpoint or model is not None:
        if model is None:
            print("\nLoading checkpoint...")
            model = load_checkpoint_model(model_path, base_model=base_model, dtype=dtype, max_batch=batch_size or 16)
        if n_mels != model.arch.n_mels:
            raise ValueError(f"--n-mels {n_mels} does not match {model.arch.name} (n_mels={model.arch.n_mels}); the reference "
                             "defaults --n-mels to 128 (large-v3), ref:scripts/evaluate_model.py:304-309")
        print("\nTranscribing test samples...")
        hypotheses = transcribe_batched(model, paths, n_mels, batch_size)
    else:
        base = load_base_model(model_path, dtype=dtype)
        decoding.require_detokenizer()
        print("\nTranscribing test samples...")
        hypotheses = [transcribe_with_model(model_path, p, is_checkpoint=False, model=base) for p in paths]

    for i in range(min(3, len(references))):                 # show first few examples
        print(f"\nSample {i + 1}:")
        print(f"  Reference:  {references[i]}")
        print(f"  Hypothesis: {hypotheses[i]}")
        print(f"  PER:  {phone_error_rate(references[i], hypotheses[i]):.2f}%")
        if pfer_available():
            print(f"  PFER: {phone_feature_error_rate(references[i], hypotheses[i]):.2f}%")

    print("\n" + "=" * 70)
    print(f"{model_name} - Overall Results")
    print("=" * 70)
    results = evaluate_batch(references, hypotheses)
    print(f"\nPER (Phone Error Rate):         {results['per']:.2f}% (±{results['per_std']:.2f}%)")
    print(f"PFER (Phone Feature Error Rate): {results['pfer']:.2f}% (±{results['pfer_std']:.2f}%)")
    print(f"Number of samples: {results['num_samples']}")
    return results


def compare_models(base_results: dict, trained_results: dict):
    """Print comparison between base and trained models (ref:scripts/evaluate_model.py:235-268)."""
    print("\n" + "=" * 70)
    print("Model Comparison")
    print("=" * 70)
    print(f"\n{'Metric':<30} {'Base Model':<15} {'Trained Model':<15} {'Improvement':<15}")
    print("-" * 70)
    per_diff = base_results["per"] - trained_results["per"]
    pfer_diff = base_results["pfer"] - trained_results["pfer"]
    print(f"{'PER (Phone Error Rate)':<30} {base_results['per']:>6.2f}%{'':<8} {trained_results['per']:>6.2f}%{'':<8} {per_diff:>+6.2f}%")
    print(f"{'PFER (Feature Error Rate)':<30} {base_results['pfer']:>6.2f}%{'':<8} {trained_results['pfer']:>6.2f}%{'':<8} {pfer_diff:>+6.2f}%")
    for bound, label in ((50, "MINIMUM VIABLE: PFER < 50% achieved!"), (30, "GOOD: PFER < 30% achieved!"),
                         (25, "EXCELLENT: PFER < 25% achieved!"), (21.2, "SOTA: Beat paper's best zero-shot result!")):
        if trained_results["pfer"] < bound:
            print(f"✅ {label}")
    return {"per_improvement": per_diff, "pfer_improvement": pfer_diff}


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description="Evaluate Whisper-IPA model")
    parser.add_argument("--checkpoint", type=str, default="checkpoints/whisper-ipa-english/checkpoint-250",
                        help="Path to trained model checkpoint")
    parser.add_argument("--base-model", type=str, default=DEFAULT_BASE, help="Base model for comparison")
    parser.add_argument("--test-data", type=str, default="data/processed/english_only_test_ipa.json", help="Path to test data JSON")
    parser.add_argument("--num-samples", type=int, default=100, help="Number of samples to evaluate (default: 100, use 0 for all)")
    parser.add_argument("--skip-base", action="store_true", help="Skip base model evaluation (only evaluate checkpoint)")
    parser.add_argument("--n-mels", type=int, default=128, help="Number of mel bins (80 for small/medium, 128 for large)")
    # additions of this implementation (the reference has no batching or dtype choice)
    parser.add_argument("--batch-size", type=int, default=64, help="utterances per GPU micro-batch")
    parser.add_argument("--dtype", type=str, default="float32", choices=["float32", "float16", "bfloat16"],
                        help="float32 = the reference's set_dtype(mx.float32); float16 = the tensor-core path")
    args = parser.parse_args(argv)
    num_samples = None if args.num_samples == 0 else args.num_samples
    if not args.skip_base:
        base_results = evaluate_model(args.base_model, args.test_data, num_samples, model_name="Base Whisper Model",
                                      is_checkpoint=False, n_mels=args.n_mels, base_model=args.base_model)
    else:
        base_results = None
    trained_results = evaluate_model(args.checkpoint, args.test_data, num_samples, model_name="Trained Checkpoint",
                                     is_checkpoint=True, n_mels=args.n_mels, base_model=args.base_model,
                                     batch_size=args.batch_size, dtype=args.dtype)
    if base_results:
        compare_models(base_results, trained_results)
    print("\n" + "=" * 70)
    print("✅ Evaluation Complete!")
    print("=" * 70)
    return trained_results


if __name__ == "__main__":
    main()
