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


def test_language_tokens_match_hf_order():
    from transformers.models.whisper.tokenization_whisper import LANGUAGES as HF_LANGUAGES
    from whisper_ipa_b200.archs import FIRST_LANGUAGE_TOKEN, LANGUAGES
    assert list(LANGUAGES) == list(HF_LANGUAGES.keys())
    small, v3 = w.ARCHS["small"], w.ARCHS["large-v3"]
    assert small.language_token("en") == 50259 == FIRST_LANGUAGE_TOKEN and small.language_token("su") == 50357
    assert small.prompt("de") == [50258, 50261, 50359, 50363]
    assert v3.language_token("yue") == 50358 and v3.prompt("yue", "translate") == [50258, 50358, 50359, 50364]
    with pytest.raises(ValueError):
        small.language_token("yue")                    # not in the pre-v3 vocabulary (50358 is <|translate|> there)
    with pytest.raises(ValueError):
        small.language_token("xx")
    assert small.language_of_token(50261) == "de"


class _FakeModel:
    """Stands in for WhisperIPA on the CPU: records which utterances each decode call saw and with which prompt."""

    def __init__(self, lang_of_utt):
        self.arch = w.ARCHS["tiny"]
        self.begin_suppress_tokens = (220, 50257)
        self._lang_of_utt = list(lang_of_utt)
        self._cached = list(range(len(lang_of_utt)))
        self.calls = []

    def encoder(self, mel):
        self._cached = list(range(mel.shape[0]))
        return torch.arange(mel.shape[0], dtype=torch.float32).view(-1, 1, 1).expand(-1, 1500, self.arch.d_model).clone()

    def set_audio_features(self, feats):
        self._cached = [int(f[0, 0]) for f in feats]

    def teacher_forced_logits(self, tokens):
        assert tokens.shape[1] == 1 and int(tokens[0, 0]) == self.arch.sot
        out = torch.zeros(tokens.shape[0], 1, self.arch.vocab)
        out[:, :, 100] = 50.0                              # a non-language token with the largest logit: must be masked out
        for b, u in enumerate(self._cached):
            out[b, 0, self.arch.language_token(self._lang_of_utt[u])] = 5.0
        return out

    def decode_tokens(self, prompt, max_new, num_beams=1, length_penalty=1.0, suppress=None, begin_suppress=None):
        self.calls.append((list(prompt), list(self._cached)))
        ids = torch.tensor([[1000 + u, prompt[1], 7] for u in self._cached], dtype=torch.int32)
        return ids, torch.full((len(self._cached),), 2, dtype=torch.int32)


def test_detect_language_and_per_language_decode():
    from whisper_ipa_b200.decoding import DecodingOptions, decode, detect_language
    mel = torch.zeros(5, 3000, 80)
    m = _FakeModel(["en", "de", "en", "ja", "de"])
    langs, probs = detect_language(m, mel)
    assert langs == ["en", "de", "en", "ja", "de"]
    assert abs(sum(probs[1].values()) - 1.0) < 1e-5 and max(probs[1], key=probs[1].get) == "de" and len(probs[1]) == 99
    res = decode(m, mel, DecodingOptions(language=None))
    assert [r.language for r in res] == ["en", "de", "en", "ja", "de"]
    # every utterance was decoded once, in the group of its language, with that language's prompt
    assert sorted(u for _, rows in m.calls for u in rows) == [0, 1, 2, 3, 4]
    for prompt, rows in m.calls:
        assert all(m.arch.language_token(["en", "de", "en", "ja", "de"][u]) == prompt[1] for u in rows)
    assert [r.tokens for r in res] == [[1000, 50259], [1001, 50261], [1002, 50259], [1003, 50266], [1004, 50261]]
    # one language for the whole batch: a single decode call on the cached encoder output, no regrouping
    m2 = _FakeModel(["fr", "fr"])
    res2 = decode(m2, torch.zeros(2, 3000, 80), DecodingOptions(language=None))
    assert len(m2.calls) == 1 and m2.calls[0][1] == [0, 1] and [r.language for r in res2] == ["fr", "fr"]
    # the reference's evaluation call (language="en") never runs the detection step
    m3 = _FakeModel(["ja"])
    r3 = decode(m3, torch.zeros(3000, 80), DecodingOptions(language="en", without_timestamps=True))
    assert r3.language == "en" and m3.calls[0][0] == [50258, 50259, 50359, 50363]
