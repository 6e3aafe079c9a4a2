"""Host-side logic of the product package (no GPU): segmentation, PER arithmetic, packing, names, sharding, audio helpers."""
import json
import os
import wave

import numpy as np
import pytest
import torch

import whisper_ipa_b200 as w
from whisper_ipa_b200 import checkpoint, metrics, parallel, pipeline


def test_tokenize_ipa_reference_assertions(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "per_cases.json"), encoding="utf-8"))
    for text, want in cases["tokenize"]:
        assert w.tokenize_ipa(text) == want
    for c in cases["corpus"]:
        assert len(w.tokenize_ipa(c["ref"])) == c["n_ref"]


def test_normalize():
    assert w.normalize_ipa_for_comparison("g a t") == "ɡat"


def test_per_from_counts_rules():
    assert metrics.per_from_counts(0, 0, 0) == 0.0
    assert metrics.per_from_counts(3, 0, 3) == 100.0
    assert metrics.per_from_counts(2, 9, 9) == (2 / 9) * 100.0
    assert metrics.per_from_counts(20, 9, 29) > 100.0


def test_summarize_matches_numpy_expression():
    s = [0.0, 33.33333333333333, 100.0, 250.0]
    out = metrics.summarize(s)
    assert out["per"] == np.mean(s) and out["per_std"] == np.std(s) and out["num_samples"] == 4


def test_pack_csr():
    flat, off = metrics._pack([[1, 2, 3], [], [7]])
    assert off.tolist() == [0, 3, 3, 4] and flat[:4].tolist() == [1, 2, 3, 7]
    flat, off = metrics._pack([[], []])
    assert off.tolist() == [0, 0, 0] and flat.size >= 1


def test_hyps_to_csr_cpu():
    ids = torch.tensor([[5, 6, 9, 9], [1, 9, 9, 9], [9, 9, 9, 9]], dtype=torch.int32)
    lens = torch.tensor([2, 1, 0], dtype=torch.int32)
    flat, off = pipeline.hyps_to_csr(ids, lens)
    assert off.tolist() == [0, 2, 3, 3] and flat.tolist() == [5, 6, 1]


def test_archs_and_prompts():
    a = w.arch_from_name("mlx-community/whisper-small-mlx")
    assert (a.d_model, a.heads, a.n_mels, a.vocab) == (768, 12, 80, 51865)
    assert a.prompt() == [50258, 50259, 50359, 50363]
    v3 = w.arch_from_name("mlx-community/whisper-large-v3-mlx")
    assert v3.n_mels == 128 and v3.prompt() == [50258, 50259, 50360, 50364]
    with pytest.raises(ValueError):
        w.arch_from_name("something-else")


def test_mlx_name_map():
    f = checkpoint.mlx_to_hf_name
    assert f("decoder.blocks.3.attn.query.weight") == "model.decoder.layers.3.self_attn.q_proj.weight"
    assert f("decoder.blocks.0.cross_attn_ln.bias") == "model.decoder.layers.0.encoder_attn_layer_norm.bias"
    assert f("decoder.blocks.11.mlp2.weight") == "model.decoder.layers.11.fc2.weight"
    assert f("decoder.token_embedding.weight") == "model.decoder.embed_tokens.weight"
    assert f("encoder.ln_post.weight") == "model.encoder.layer_norm.weight"
    assert f("model.decoder.layer_norm.bias") == "model.decoder.layer_norm.bias"
    sd = checkpoint.to_hf_state_dict({"encoder.conv1.weight": np.zeros((8, 3, 5), np.float32)}, w.ARCHS["tiny"])
    assert tuple(sd["model.encoder.conv1.weight"].shape) == (8, 5, 3)
    with pytest.raises(KeyError):
        f("decoder.unknown.weight")


def test_shard_indices_partition():
    for n, ws in ((10, 4), (7, 8), (0, 2), (16, 2)):
        seen = sorted(i for r in range(ws) for i in parallel.shard_indices(n, r, ws))
        assert seen == list(range(n))


def test_pad_or_trim_and_load_audio(tmp_path):
    x = np.arange(10, dtype=np.float32)
    assert w.pad_or_trim(x, 16).shape == (16,) and w.pad_or_trim(x, 16)[10:].sum() == 0
    assert w.pad_or_trim(x, 4).tolist() == [0, 1, 2, 3]
    t = torch.arange(10.0)[None]
    assert tuple(w.pad_or_trim(t, 16).shape) == (1, 16) and tuple(w.pad_or_trim(t, 4).shape) == (1, 4)
    p = tmp_path / "a.wav"
    pcm = (np.sin(np.arange(8000) / 10.0) * 20000).astype("<i2")
    with wave.open(str(p), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(8000); f.writeframes(pcm.tobytes())
    a = w.load_audio(str(p))
    assert a.dtype == np.float32 and abs(len(a) - 16000) <= 1 and np.abs(a).max() <= 1.0


def test_decoding_options_defaults():
    o = w.DecodingOptions(language="en", without_timestamps=True)
    assert o.temperature == 0.0 and o.beam_size is None and o.sample_len is None


def test_no_gpu_is_loud():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        w.WhisperIPA("tiny")
