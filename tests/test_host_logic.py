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


# ---- round 2: suppress lists, checkpoints, detokenizer, resampler plan, long-form rules ----------------------------------
def test_default_suppress_tokens_follow_the_tokenizer_lists():
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS_MULTI
    from whisper_ipa_b200.archs import NON_SPEECH_TOKENS
    assert list(NON_SPEECH_TOKENS) == [t for t in NON_SPEECH_TOKENS_MULTI if t < 50257]
    small, v3 = w.ARCHS["small"], w.ARCHS["large-v3"]
    # "-1" = non-speech symbols + transcribe, translate, sot, sot_prev, sot_lm, no_speech (openai/whisper-small's
    # generation_config.suppress_tokens is exactly NON_SPEECH_TOKENS_MULTI + {50358, 50359})
    assert small.default_suppress_tokens() == sorted(set(NON_SPEECH_TOKENS_MULTI) | {50358, 50359})
    assert set(v3.default_suppress_tokens()) - set(NON_SPEECH_TOKENS) == {50258, 50359, 50360, 50361, 50362, 50363}
    assert small.resolve_suppress_tokens(None) == [] and small.resolve_suppress_tokens("") == []
    assert small.resolve_suppress_tokens("-1") == small.default_suppress_tokens()
    assert small.resolve_suppress_tokens("5,7") == [5, 7]
    assert small.resolve_suppress_tokens([-1, 123]) == sorted(set(small.default_suppress_tokens()) | {123})
    with pytest.raises(ValueError):
        small.resolve_suppress_tokens([60000])
    assert (small.no_timestamps, small.timestamp_begin, v3.no_timestamps, v3.timestamp_begin) == (50363, 50364, 50364, 50365)


def test_decode_passes_reference_defaults_to_the_engine():
    """DecodingOptions(language="en", without_timestamps=True): suppress_tokens="-1" expands to the default list, 224 tokens
    are sampled AFTER the prompt (mlx_whisper's sample_len), blank + EOT are masked at the first step."""
    from whisper_ipa_b200.decoding import DecodingOptions, decode
    seen = {}

    class M(_FakeModel):
        def decode_tokens(self, prompt, max_new, num_beams=1, length_penalty=1.0, suppress=None, begin_suppress=None):
            seen.update(max_new=max_new, suppress=list(suppress), begin=list(begin_suppress))
            return super().decode_tokens(prompt, max_new)
    m = M(["en"])
    decode(m, torch.zeros(1, 3000, 80), DecodingOptions(language="en", without_timestamps=True))
    assert seen["max_new"] == 224 and seen["suppress"] == m.arch.default_suppress_tokens() and seen["begin"] == [220, 50257]
    decode(m, torch.zeros(1, 3000, 80), DecodingOptions(language="en", suppress_tokens=None, sample_len=500, suppress_blank=False))
    assert seen["max_new"] == 444 and seen["suppress"] == [] and seen["begin"] == [50257]


def test_mlx_checkpoint_round_trip_with_non_weight_entries(tiny_sd, tmp_path):
    """A full MLX-named key set, as the reference saves it (flatten_params(model.parameters()),
    ref:scripts/train_whisper_ipa.py:420-422: it includes `alignment_heads` and has no encoder position table), maps back to
    the HF names and values; the `decoder.` prefix filter is applied before any name is mapped."""
    from safetensors.torch import save_file
    mlx = checkpoint.hf_to_mlx_state_dict(tiny_sd)
    assert "encoder.blocks.0.attn.query.weight" in mlx and "decoder.blocks.3.mlp2.bias" in mlx
    assert tuple(mlx["encoder.conv1.weight"].shape) == (384, 3, 80)            # MLX conv layout [out, k, in]
    mlx["alignment_heads"] = torch.zeros(6, 2)
    mlx["encoder.positional_embedding"] = torch.zeros(1500, 384)
    back = checkpoint.to_hf_state_dict(mlx, w.ARCHS["tiny"])
    assert set(tiny_sd) - set(back) == {"model.encoder.embed_positions.weight", "proj_out.weight"}
    assert all(torch.equal(back[k], tiny_sd[k]) for k in back)
    save_file({k: v.contiguous() for k, v in mlx.items()}, str(tmp_path / "model.safetensors"))
    dec = checkpoint.load_weights_dir(str(tmp_path), w.ARCHS["tiny"], prefix="decoder.")
    assert dec and all(k.startswith("model.decoder.") for k in dec)
    assert len(dec) == sum(1 for k in mlx if k.startswith("decoder."))
    with pytest.raises(KeyError):
        checkpoint.to_hf_state_dict({"decoder.blocks.0.nonsense.weight": torch.zeros(1)}, w.ARCHS["tiny"])


def test_detokenizer_is_required_for_text(tmp_path):
    import base64
    from whisper_ipa_b200 import decoding
    decoding.set_detokenizer(None)
    r = decoding.DecodingResult(None, "en", [3, 1, 50257])
    assert r.tokens == [3, 1, 50257]
    with pytest.raises(RuntimeError, match="no detokenizer"):
        r.text
    with pytest.raises(RuntimeError, match="no detokenizer"):
        w.transcribe_batched(None, [], 80)                     # refused before any work, never swallowed per sample
    decoding.set_detokenizer("ids")
    assert decoding.DecodingResult(None, "en", [3, 1]).text == "3 1"
    # tiktoken-format vocabulary file, as mlx_whisper ships it: base64(token bytes) + rank per line
    toks = ["k".encode(), "æ".encode(), "t".encode(), " ".encode(), "ʃ".encode()[:1], "ʃ".encode()[1:]]
    with open(tmp_path / "multilingual.tiktoken", "wb") as f:
        for i, t in enumerate(toks):
            f.write(base64.b64encode(t) + b" " + str(i).encode() + b"\n")
    fn = decoding.load_detokenizer(str(tmp_path))
    assert fn([0, 1, 2]) == "kæt" and fn([4, 5, 3, 0, 50257, 50363]) == "ʃ k"      # bytes joined across tokens, specials dropped
    assert decoding.DecodingResult(None, "en", [0, 1, 2, 50257]).text == "kæt"
    decoding.set_detokenizer(None)
    with pytest.raises(FileNotFoundError):
        decoding.load_detokenizer(str(tmp_path / "nowhere"))


def test_resample_plan_matches_scipy():
    """The filter and the index bookkeeping handed to the GPU resampler reproduce scipy.signal.resample_poly (numpy
    emulation of the kernel's formula; the kernel itself is compared with scipy in the GPU tests)."""
    from scipy.signal import firwin, resample_poly
    from whisper_ipa_b200.ingest import resample_plan
    rng = np.random.default_rng(0)
    for rate in (48000, 44100, 8000, 22050, 16000):
        plan = resample_plan(rate)
        x = rng.standard_normal(1200).astype(np.float32)
        ref = resample_poly(x, plan.up, plan.down) if rate != 16000 else x
        if rate != 16000:
            m = max(plan.up, plan.down)
            assert np.abs(plan.taps - (firwin(20 * m + 1, 1.0 / m, window=("kaiser", 5.0)) * plan.up)).max() < 1e-6
        nt = len(plan.taps)
        got = np.zeros(len(ref))
        for n in range(len(ref)):
            c = n * plan.down + plan.c0
            lo = max(0, -(-(c - nt + 1) // plan.up))
            hi = min(len(x) - 1, c // plan.up)
            i = np.arange(lo, hi + 1)
            got[n] = np.dot(plan.taps[c - i * plan.up].astype(np.float64), x[i])
        assert np.abs(got - ref).max() < 2e-6, rate
        assert plan.frames_needed(len(ref)) >= min(len(x), (len(ref) - 1) * plan.down // plan.up)


def test_timestamp_rules_match_hf_processor():
    """transcribe.apply_timestamp_rules (the long-form branch's logit filter) against HF's WhisperTimeStampLogitsProcessor
    on random logits and random sampled prefixes (HF:generation/logits_process.py:1995-2043)."""
    from types import SimpleNamespace
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor
    from whisper_ipa_b200.transcribe import apply_timestamp_rules
    arch = w.ARCHS["tiny"]
    cfg = SimpleNamespace(no_timestamps_token_id=arch.no_timestamps, eos_token_id=arch.eot, bos_token_id=arch.eot,
                          max_initial_timestamp_index=50, _detect_timestamp_from_logprob=True)
    g = torch.Generator().manual_seed(0)
    tb = arch.timestamp_begin
    for trial in range(40):
        n = int(torch.randint(0, 6, (1,), generator=g))
        seq = []
        for _ in range(n):
            seq.append(int(torch.randint(tb, tb + 200, (1,), generator=g)) if torch.rand(1, generator=g) < 0.5
                       else int(torch.randint(0, 50000, (1,), generator=g)))
        logits = torch.randn(arch.vocab, generator=g) * (1.0 if trial % 2 else 6.0)
        prompt = [arch.sot, 50259, arch.transcribe]
        hf_proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))
        want = hf_proc(torch.tensor([prompt + seq]), logits[None].clone())[0]
        got = apply_timestamp_rules(logits.clone(), seq, arch, first=(n == 0))
        assert torch.equal(torch.isinf(got), torch.isinf(want)) and torch.equal(got[~torch.isinf(got)], want[~torch.isinf(want)]), seq


def test_panphon_csv_table_loader(tmp_path):
    from whisper_ipa_b200 import metrics
    head = "ipa,syl,son,cons,cont,delrel,lat,nas,strid,voi,sg,cg,ant,cor,distr,lab,hi,lo,back,round,velaric,tense,long,hitone,hireg"
    rows = ["p," + ",".join(["-"] * 24), "a," + ",".join(["+"] * 24), "pʰ," + ",".join(["0"] * 23 + ["+"])]
    (tmp_path / "ipa_all.csv").write_text("\n".join([head] + rows) + "\n", encoding="utf-8")
    try:
        table = metrics.load_feature_table(str(tmp_path / "ipa_all.csv"))
        assert table["p"] == [-1] * 24 and table["a"] == [1] * 24
        assert metrics._phone_features("pʰ").tolist() == [0] * 23 + [1]         # longest segment wins
        assert metrics._phone_features("xa").tolist() == [1] * 24                # first KNOWN segment, as panphon's finditer
        assert metrics._phone_features("x").tolist() == [0] * 24                 # unknown phone -> zero vector
        assert metrics.pfer_available()
    finally:
        metrics.set_feature_table(None)


def test_compare_models_prints_and_returns(capsys):
    out = w.compare_models({"per": 60.0, "pfer": 40.0}, {"per": 30.0, "pfer": 24.0})
    text = capsys.readouterr().out
    assert out == {"per_improvement": 30.0, "pfer_improvement": 16.0}
    assert "Model Comparison" in text and "+30.00%" in text and "EXCELLENT" in text and "SOTA" not in text


def test_cli_flags_match_the_reference():
    import importlib
    em = importlib.import_module("whisper_ipa_b200.evaluate_model")     # the package attribute of that name is the function
    seen = {}

    def fake_eval(model_path, test_data_path, num_samples=None, **kw):
        seen.setdefault("calls", []).append((model_path, test_data_path, num_samples, kw))
        return {"per": 1.0, "pfer": 2.0, "per_std": 0.0, "pfer_std": 0.0, "num_samples": 0}
    old = em.evaluate_model
    em.evaluate_model = fake_eval
    try:
        em.main(["--checkpoint", "ck", "--base-model", "whisper-small", "--test-data", "t.json", "--num-samples", "0", "--n-mels", "80"])
        assert [c[0] for c in seen["calls"]] == ["whisper-small", "ck"] and seen["calls"][1][2] is None
        assert seen["calls"][0][3]["is_checkpoint"] is False and seen["calls"][1][3]["is_checkpoint"] is True
        seen.clear()
        em.main(["--checkpoint", "ck", "--skip-base"])
        assert len(seen["calls"]) == 1 and seen["calls"][0][2] == 100 and seen["calls"][0][3]["n_mels"] == 128
    finally:
        em.evaluate_model = old


# ---- long-form windowing on the CPU: a scripted model stands in for the GPU ---------------------------------------------------
class _ScriptedModel:
    """decode_begin / decode_next return logits that make a fixed token script the argmax (timestamp rules permitting); the
    script is chosen per window from the number of windows decoded so far."""

    def __init__(self, scripts, no_speech=0.0):
        self.arch = w.ARCHS["tiny"]
        self.device = torch.device("cpu")
        self.scripts = scripts
        self.no_speech = no_speech
        self.window = -1
        self.prompts = []
        self.encoded = []

    def encoder(self, mel, return_features=True):
        self.encoded.append(tuple(mel.shape))
        return None

    def _logits(self, k):
        out = torch.full((1, self.arch.vocab), -20.0)
        script = self.scripts[min(self.window, len(self.scripts) - 1)]
        out[0, script[k] if k < len(script) else self.arch.eot] = 10.0
        return out

    def decode_begin(self, tokens):
        toks = list(tokens[0])
        if toks[-1] == self.arch.sot and len(toks) == 1 or toks[-1] == self.arch.sot:      # the no-speech probe
            out = torch.full((1, self.arch.vocab), -20.0)
            out[0, self.arch.no_speech] = float(np.log(max(self.no_speech, 1e-9) / max(1 - self.no_speech, 1e-9))) - 20.0 + 20.0
            out[0, 0] = 0.0
            return out
        self.window += 1
        self.prompts.append(toks)
        self.k = 0
        return self._logits(0)

    def decode_next(self, tok):
        self.k += 1
        return self._logits(self.k)


def test_long_form_windows_follow_the_timestamps(monkeypatch):
    """transcribe(): the seek advances to the last timestamp of a window that ends in a closed pair, the previous window's text
    conditions the next one behind <|startofprev|>, a window whose no-speech probability is high AND whose log-probability is
    low is skipped, and the text is the concatenation of the segments' text tokens."""
    from whisper_ipa_b200 import decoding
    from whisper_ipa_b200 import transcribe as tr
    arch = w.ARCHS["tiny"]
    tb = arch.timestamp_begin
    monkeypatch.setattr(tr, "log_mel_features", lambda clip, n_mels: torch.zeros(1, n_mels, 3000))
    decoding.set_detokenizer("ids")
    try:
        # window 0: <0.00> 11 12 <10.00><10.00> 13 <20.00><20.00>  -> two segments, seek moves to 20 s (1000 positions * 2 frames)
        # window 1 (starts at 20 s): <0.00> 21 <5.00> then EOT       -> single closing timestamp: whole window consumed
        s0 = [tb, 11, 12, tb + 500, tb + 500, 13, tb + 1000, tb + 1000]
        s1 = [tb, 21, tb + 250]
        m = _ScriptedModel([s0, s1])
        audio = np.zeros(16000 * 45, np.float32)                       # 45 s
        out = tr.transcribe(audio, m, language="en", temperature=0.0, compression_ratio_threshold=None, logprob_threshold=None,
                            no_speech_threshold=None)
        segs = out["segments"]
        assert [(round(s["start"], 2), round(s["end"], 2)) for s in segs] == [(0.0, 10.0), (10.0, 20.0), (20.0, 25.0)]
        assert [s["seek"] for s in segs] == [0, 0, 2000]
        assert out["text"] == "11 12 13 21"
        # the second window was conditioned on the first one's tokens: <|startofprev|> + previous tokens + sot sequence
        assert m.prompts[0] == [arch.sot, arch.language_token("en"), arch.transcribe]
        # (the trailing <20.00> opens the next segment and belongs to no slice, as in the reference algorithm)
        assert m.prompts[1][0] == arch.sot_prev and m.prompts[1][1:-3] == s0[:-1] and m.prompts[1][-3:] == m.prompts[0]
        assert len(m.encoded) == 2
        # silence: P(no speech) high and the average log-probability below the threshold -> every window skipped, empty text
        m2 = _ScriptedModel([[tb, 11, tb + 100]], no_speech=0.99)
        out2 = tr.transcribe(np.zeros(16000 * 10, np.float32), m2, language="en", temperature=0.0, compression_ratio_threshold=None,
                             logprob_threshold=0.5, no_speech_threshold=0.6)
        assert out2["segments"] == [] and out2["text"] == ""
        # a confident window survives the same no-speech probability (avg_logprob above the threshold)
        m3 = _ScriptedModel([[tb, 11, tb + 100]], no_speech=0.99)
        out3 = tr.transcribe(np.zeros(16000 * 10, np.float32), m3, language="en", temperature=0.0, compression_ratio_threshold=None,
                             logprob_threshold=-1.0, no_speech_threshold=0.6)
        assert out3["text"] == "11"
    finally:
        decoding.set_detokenizer(None)


def test_read_pcm_host_side(tmp_path):
    """ingest.read_pcm (the worker-thread half of the audio ingest, no GPU involved): PCM16 WAVs come back as raw interleaved
    samples cut to what 30 s of output can need, other sample widths as host-decoded float audio, unreadable files raise."""
    from whisper_ipa_b200.ingest import read_pcm, resample_plan
    pcm = (np.arange(48000 * 2 * 2) % 2000 - 1000).astype("<i2").reshape(-1, 2)          # 2 s of 48 kHz stereo
    p = tmp_path / "s.wav"
    with wave.open(str(p), "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(48000); f.writeframes(pcm.tobytes())
    it = read_pcm(str(p))
    assert it["rate"] == 48000 and it["ch"] == 2 and it["frames"] == 96000 and np.array_equal(it["pcm"], pcm.reshape(-1))
    long = np.zeros((48000 * 40, 1), "<i2")                                              # 40 s mono: only ~30 s are read
    p2 = tmp_path / "l.wav"
    with wave.open(str(p2), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(48000); f.writeframes(long.tobytes())
    it2 = read_pcm(str(p2))
    assert it2["frames"] == resample_plan(48000).frames_needed(480000) < 48000 * 31
    p3 = tmp_path / "e.wav"
    with wave.open(str(p3), "wb") as f:
        f.setnchannels(1); f.setsampwidth(1); f.setframerate(16000); f.writeframes(bytes(range(256)) * 10)
    it3 = read_pcm(str(p3))
    assert "f32" in it3 and it3["f32"].dtype == np.float32 and len(it3["f32"]) == 2560
    with pytest.raises(Exception):
        read_pcm(str(tmp_path / "missing.wav"))
