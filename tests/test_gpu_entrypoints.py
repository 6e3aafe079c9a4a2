"""GPU tests of the reference-facing entry points and the pieces either side of the hot path (SURVEY.md §8 rows a1, a5, a6,
f2, f3, f4): audio ingest (resampler kernel + worker pool), evaluate_model / load_checkpoint_model / transcribe_file on
real files, language detection + per-language decode (the training-time validate() caller), stepwise decoding with
logits processors, and the long-form base-model branch."""
import json
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w(built_lib):
    import whisper_ipa_b200 as w
    torch.backends.cuda.matmul.allow_tf32 = False
    return w


def _write_wav(path, pcm, rate):
    """pcm int16 [frames, channels]"""
    with wave.open(str(path), "wb") as f:
        f.setnchannels(pcm.shape[1]); f.setsampwidth(2); f.setframerate(rate); f.writeframes(pcm.astype("<i2").tobytes())


def _host_reference_audio(pcm, rate):
    """What audio.load_audio + pad_or_trim give on the host: / 32768 -> channel mean -> scipy resample_poly -> 30 s."""
    from scipy.signal import resample_poly
    x = pcm.astype(np.float32) / 32768.0
    x = x.mean(axis=1) if pcm.shape[1] > 1 else x[:, 0]
    if rate != 16000:
        g = np.gcd(rate, 16000)
        x = resample_poly(x, 16000 // g, rate // g).astype(np.float32)
    out = np.zeros(480000, np.float32)
    out[:min(len(x), 480000)] = x[:480000]
    return out


def _pcm(rng, seconds, rate, ch):
    n = int(seconds * rate)
    t = np.arange(n)[:, None] / rate
    sig = 0.3 * np.sin(2 * np.pi * (220.0 + 110.0 * np.arange(ch)[None, :]) * t) + 0.1 * rng.standard_normal((n, ch))
    return np.clip(sig * 32768.0, -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("rate,ch", [(48000, 2), (44100, 1), (8000, 1), (16000, 2), (22050, 2), (32000, 1)])
def test_resample_kernel_matches_scipy(w, rate, ch):
    """wipa_resample_pcm16 vs the host path (scipy.signal.resample_poly): ragged clip lengths in one launch, a clip longer
    than 30 s (cut), one of a few samples, stereo down-mix."""
    from whisper_ipa_b200 import _lib
    from whisper_ipa_b200.ingest import resample_plan
    rng = np.random.default_rng(rate + ch)
    clips = [_pcm(rng, s, rate, ch) for s in (0.31, 4.0, 31.0, 0.002)]
    plan = resample_plan(rate)
    flat = np.concatenate([c.reshape(-1) for c in clips])
    offs = np.cumsum([0] + [c.size for c in clips[:-1]]).astype(np.int64)
    frames = np.asarray([c.shape[0] for c in clips], np.int32)
    d_pcm, d_off, d_fr = (torch.from_numpy(a).cuda() for a in (flat, offs, frames))
    taps = torch.from_numpy(plan.taps).cuda()
    out = torch.full((len(clips), 480000), float("nan"), device="cuda")
    _lib.check(_lib.lib().wipa_resample_pcm16(d_pcm.data_ptr(), d_off.data_ptr(), d_fr.data_ptr(), len(clips), ch, plan.up, plan.down,
                                              taps.data_ptr(), taps.numel(), plan.c0, out.data_ptr(), 480000,
                                              torch.cuda.current_stream().cuda_stream), "wipa_resample_pcm16")
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    for i, c in enumerate(clips):
        want = _host_reference_audio(c, rate)
        err = np.abs(got[i] - want).max()
        assert err < 3e-6, f"clip {i}: max err {err}"


def test_audio_ingest_pool_mixed_files(w, tmp_path):
    """AudioIngest.load_batch over a mixed bag of files: different rates / channel counts in one batch (several kernel
    groups, non-contiguous rows), an 8-bit WAV (host path), a missing file (silence + the exception) - every row must equal
    pad_or_trim(load_audio(path)) of the host path."""
    from whisper_ipa_b200.ingest import AudioIngest
    rng = np.random.default_rng(7)
    specs = [(48000, 2, 2.0), (16000, 1, 1.0), (48000, 2, 3.5), (44100, 1, 0.5), (16000, 1, 30.5), (48000, 1, 1.0)]
    paths = []
    for i, (rate, ch, sec) in enumerate(specs):
        p = tmp_path / f"a{i}.wav"
        _write_wav(p, _pcm(rng, sec, rate, ch), rate)
        paths.append(str(p))
    p8 = tmp_path / "eight.wav"
    with wave.open(str(p8), "wb") as f:
        f.setnchannels(1); f.setsampwidth(1); f.setframerate(16000)
        f.writeframes((rng.integers(0, 255, 8000)).astype(np.uint8).tobytes())
    paths.insert(3, str(p8))
    paths.insert(5, str(tmp_path / "missing.wav"))
    ing = AudioIngest(chunk=3)
    try:
        audio, errors = ing.load_batch(paths)
        torch.cuda.synchronize()
        got = audio.cpu().numpy()
        assert got.shape == (len(paths), 480000)
        for i, p in enumerate(paths):
            if p.endswith("missing.wav"):
                assert errors[i] is not None and not got[i].any()
                continue
            assert errors[i] is None
            want = w.pad_or_trim(w.load_audio(p))
            assert np.abs(got[i] - want).max() < 3e-6, p
        # the prefetching iterator yields the same rows
        rows = {}
        for idx, a, errs in ing.iter_batches(paths, 3):
            for j, i in enumerate(idx):
                rows[i] = a[j].cpu().numpy()
        assert sorted(rows) == list(range(len(paths))) and all(np.array_equal(rows[i], got[i]) for i in rows)
    finally:
        ing.close()


def _ipa_detok(ids):
    """A stand-in vocabulary for random-init models: every text token is one letter of the IPA Extensions block."""
    return "".join(chr(0x250 + int(i) % 96) for i in ids if int(i) < 50257)


def test_evaluate_model_end_to_end_matches_hf_and_per_oracle(w, tiny_sd, tmp_path, golden_dir):
    """evaluate_model(...) on a JSON of short 48 kHz stereo WAVs with random-init weights: per / per_std / per_scores must
    equal HF generate (same prompt, suppress lists, 224 sampled tokens) + the CPU PER oracle on the same files."""
    from oracle import hf_reference as hf
    from oracle import per_oracle as po
    from whisper_ipa_b200 import decoding
    rng = np.random.default_rng(11)
    corpus = json.load(open(os.path.join(golden_dir, "per_cases.json"), encoding="utf-8"))["corpus"]
    samples = []
    for i, sec in enumerate((1.5, 2.0, 3.0, 2.5, 1.0)):
        p = tmp_path / f"utt{i}.wav"
        _write_wav(p, _pcm(rng, sec, 48000, 2), 48000)
        samples.append({"audio_path": str(p), "ipa_transcription": corpus[i]["ref"]})
    samples.append({"audio_path": str(tmp_path / "gone.wav"), "ipa_transcription": corpus[5]["ref"]})     # -> "" hypothesis
    test_json = tmp_path / "test.json"
    test_json.write_text(json.dumps(samples, ensure_ascii=False), encoding="utf-8")

    m = w.WhisperIPA("tiny", dtype="float32", max_batch=4)
    m.load_state_dict(tiny_sd)
    decoding.set_detokenizer(_ipa_detok)
    try:
        res = w.evaluate_model("unused", str(test_json), model_name="tiny random-init", is_checkpoint=True, n_mels=80, model=m,
                               batch_size=4)
    finally:
        decoding.set_detokenizer(None)
        m.close()
    arch = w.ARCHS["tiny"]
    hf_model = hf.build_hf_model("tiny", seed=0)
    audio = np.stack([w.pad_or_trim(w.load_audio(s["audio_path"])) for s in samples[:-1]])
    prompt = arch.prompt("en", "transcribe", True)
    import warnings
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ids = hf_model.generate(hf.hf_log_mel(audio, 80), decoder_input_ids=torch.tensor([prompt] * len(audio)), max_new_tokens=224,
                                do_sample=False, suppress_tokens=arch.default_suppress_tokens())
    hyps = [_ipa_detok(r).strip() for r in ids.tolist()] + [""]
    want = [po.phone_error_rate(s["ipa_transcription"], h) for s, h in zip(samples, hyps)]
    assert res["num_samples"] == len(samples)
    assert res["per_scores"] == want, (res["per_scores"], want)
    assert res["per"] == np.mean(want) and res["per_std"] == np.std(want)
    assert res["per_scores"][-1] == 100.0                      # unreadable file -> empty hypothesis, as the reference


def test_mlx_named_checkpoint_loads_into_the_gpu_model(w, tiny_sd, tiny_gain_sd, tmp_path):
    """load_checkpoint_model: MLX-named base weights (npz, NLC conv layout, no encoder position table) + a checkpoint whose
    model.safetensors carries `alignment_heads`, trained decoder.* tensors and (to be ignored) encoder tensors; the ids must
    equal those of a model loaded from the equivalent HF-named dict (ref:scripts/evaluate_model.py:41-75)."""
    from safetensors.torch import save_file
    from whisper_ipa_b200 import checkpoint
    from oracle import whisper_oracle as wo
    base_dir = tmp_path / "whisper-tiny-mlx"
    base_dir.mkdir()
    np.savez(base_dir / "weights.npz", **{k: v.numpy() for k, v in checkpoint.hf_to_mlx_state_dict(tiny_sd).items()})
    ck_dir = tmp_path / "checkpoint-8000"
    ck_dir.mkdir()
    trained = checkpoint.hf_to_mlx_state_dict(tiny_gain_sd)          # "trained" = a different decoder AND a different encoder
    trained["alignment_heads"] = torch.zeros(6, 2)
    save_file({k: v.contiguous() for k, v in trained.items()}, str(ck_dir / "model.safetensors"))

    m = w.load_checkpoint_model(str(ck_dir), base_model=str(base_dir), max_batch=2)
    merged = {k: (tiny_gain_sd[k] if k.startswith("model.decoder.") or k == "proj_out.weight" else v) for k, v in tiny_sd.items()}
    ref = w.WhisperIPA("tiny", dtype="float32", max_batch=2)
    ref.load_state_dict(merged)
    audio = wo.synthetic_audio(2)
    mel = w.log_mel_features(audio, 80)
    prompt = torch.tensor([wo.PROMPT_PRE_V3] * 2)
    enc_a, enc_b = m.encoder(mel).cpu(), ref.encoder(mel).cpu()
    # the MLX files carry no encoder position table: the model's own construction of it must be HF's, bit for bit
    assert torch.equal(enc_a, enc_b)
    a = m.generate(mel, decoder_input_ids=prompt, max_new_tokens=24).cpu()
    b = ref.generate(mel, decoder_input_ids=prompt, max_new_tokens=24).cpu()
    assert torch.equal(a, b)
    base_only = w.load_checkpoint_model(str(tmp_path / "no-such-checkpoint"), base_model=str(base_dir), max_batch=2)   # WARNING path
    c = base_only.generate(mel, decoder_input_ids=prompt, max_new_tokens=8).cpu()
    assert c.shape == (2, 8)
    from whisper_ipa_b200 import transcribe_single
    with pytest.raises(SystemExit):                                      # transcribe_single's loader exits instead
        transcribe_single.load_checkpoint_model(str(tmp_path / "no-such-checkpoint"), base_model=str(base_dir))
    for x in (m, ref, base_only):
        x.close()


def test_language_none_matches_hf_detect_language_and_per_row_prompts(w, tiny_gain_sd):
    """The training-time validate() caller (ref:scripts/train_whisper_ipa.py:334-356): batch of 4 mels,
    DecodingOptions(language=None, without_timestamps=True) -> language detection (one decoder step on <|sot|>), then a decode
    with each row's own <|lang|> prompt.  Languages vs HF's detect_language, tokens vs HF generate with per-row prompts."""
    import warnings
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    from whisper_ipa_b200.archs import LANGUAGES
    arch = w.ARCHS["tiny"]
    hf_model = hf.build_hf_model("tiny", seed=0, init_gain=3.0)
    gc = hf_model.generation_config
    gc.lang_to_id = {f"<|{l}|>": 50259 + i for i, l in enumerate(LANGUAGES[:99])}
    gc.task_to_id = {"transcribe": 50359, "translate": 50358}
    gc.is_multilingual = True
    gc.no_timestamps_token_id = 50363
    # four clips whose detected languages differ: scale the noise so the encoder output differs materially
    audio = wo.synthetic_audio(4) * np.asarray([1.0, 0.02, 3.0, 0.3], np.float32)[:, None]
    feats_hf = hf.hf_log_mel(audio, 80)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_lang = [arch.language_of_token(t) for t in hf_model.detect_language(input_features=feats_hf).tolist()]
        prompts = torch.tensor([arch.prompt(l, "transcribe", True) for l in want_lang])
        want_ids = hf_model.generate(feats_hf, decoder_input_ids=prompts, max_new_tokens=20, do_sample=False,
                                     suppress_tokens=arch.default_suppress_tokens())
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=4)
    m.load_state_dict(tiny_gain_sd)
    mel = w.log_mel_spectrogram(audio, n_mels=80)                     # reference layout [B, 3000, n_mels]
    langs, probs = w.detect_language(m, mel)
    assert langs == want_lang
    assert all(abs(sum(p.values()) - 1.0) < 1e-4 for p in probs)
    res = m.decode(mel, w.DecodingOptions(language=None, without_timestamps=True, fp16=False, sample_len=20))
    assert [r.language for r in res] == want_lang
    # HF's generate() is its long-form loop: a row that samples timestamp tokens (nothing suppresses them here) gets a second
    # segment appended, so only the first window's 20 tokens are comparable
    got = torch.tensor([r.tokens + [arch.eot] * (20 - len(r.tokens)) for r in res])
    assert torch.equal(got, want_ids[:, :20])
    print(f"\n[language=None] detected {want_lang}")
    m.close()


@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_stepwise_decode_equals_teacher_forced_logits(w, tiny_gain_sd, dtype):
    """wipa_decode_begin / wipa_decode_next hand back the logits of the very kernels the fused decode runs: they must be
    bit-identical to wipa_decode_logits on the same tokens, with per-row prompts."""
    from oracle import whisper_oracle as wo
    m = w.WhisperIPA("tiny", dtype=dtype, max_batch=3)
    m.load_state_dict(tiny_gain_sd)
    m.encoder(w.log_mel_features(wo.synthetic_audio(3), 80), return_features=False)
    g = torch.Generator().manual_seed(5)
    toks = torch.randint(0, 50000, (3, 9), generator=g)
    toks[:, 0] = 50258
    want = m.teacher_forced_logits(toks)
    got = [m.decode_begin(toks[:, :4])]
    for t in range(4, 9):
        got.append(m.decode_next(toks[:, t].cuda()))
    got = torch.stack(got, dim=1)
    assert torch.equal(got, want[:, 3:])
    m.close()


def test_generate_with_logits_processor_matches_hf(w, tiny_gain_sd):
    """HF face: generate(..., logits_processor=[...]) - a processor written against HF's interface gives HF's own ids."""
    import warnings
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo

    class Bias:                                        # favours even token ids; stateless, device-agnostic
        def __call__(self, input_ids, scores):
            s = scores.clone()
            s[:, 1::2] -= 0.05 * (input_ids.shape[1] % 3)
            return s
    audio = wo.synthetic_audio(2)
    hf_model = hf.build_hf_model("tiny", seed=0, init_gain=3.0)
    prompt = torch.tensor([wo.PROMPT_PRE_V3] * 2)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = hf_model.generate(hf.hf_log_mel(audio, 80), decoder_input_ids=prompt, max_new_tokens=12, do_sample=False,
                                 logits_processor=[Bias()])
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=2)
    m.load_state_dict(tiny_gain_sd)
    got = m.generate(w.log_mel_features(audio, 80), decoder_input_ids=prompt, max_new_tokens=12, logits_processor=[Bias()]).cpu()
    m.close()
    assert torch.equal(got, want)


def test_long_form_first_window_and_sliding(w, tiny_gain_sd):
    """Base-model branch (mlx_whisper.transcribe, ref:scripts/evaluate_model.py:112-119): the first window's tokens, decoded
    with timestamp rules + suppress lists at temperature 0, must equal the CPU oracle run through the HF-pinned rules; a
    40 s file is consumed by sliding windows (seek follows the timestamps) and yields ordered segments."""
    from oracle import whisper_oracle as wo
    from whisper_ipa_b200 import decoding
    from whisper_ipa_b200 import transcribe as tr
    arch = w.ARCHS["tiny"]
    sd = tiny_gain_sd
    dims = wo.Dims.from_arch("tiny")
    audio = wo.synthetic_audio(1)
    enc = wo.encoder_forward(sd, dims, wo.log_mel_spectrogram(torch.from_numpy(audio), 80))
    xkv = wo.cross_kv(sd, dims, enc)
    prompt = [arch.sot, arch.language_token("en"), arch.transcribe]
    toks, seq = torch.tensor([prompt]), []
    sup = arch.default_suppress_tokens()
    for i in range(16):
        logits = wo.decoder_forward(sd, dims, toks, 0, xkv, [None] * dims.n_dec)[0, -1].float().clone()
        if i == 0:
            logits[[220, arch.eot]] = -float("inf")
        logits[sup] = -float("inf")
        logits = tr.apply_timestamp_rules(logits, seq, arch, first=(i == 0))
        seq.append(int(logits.argmax()))
        toks = torch.cat([toks, torch.tensor([[seq[-1]]])], dim=1)
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=1)
    m.load_state_dict(sd)
    decoding.set_detokenizer("ids")
    try:
        m.encoder(w.log_mel_features(audio, 80), return_features=False)
        win = tr.decode_window(m, [], "en", "transcribe", 0.0, sample_len=16)
        assert win.tokens == seq
        assert 0.0 <= win.no_speech_prob <= 1.0 and win.avg_logprob < 0.0
        long_audio = np.concatenate([audio[0], audio[0][:160000]])                   # 40 s
        out = tr.transcribe(long_audio, m, language="en", compression_ratio_threshold=None, logprob_threshold=None,
                            no_speech_threshold=None, temperature=0.0)
        segs = out["segments"]
        assert out["language"] == "en" and len(segs) >= 2
        assert all(s["end"] >= s["start"] >= 0.0 for s in segs)
        assert [s["start"] for s in segs] == sorted(s["start"] for s in segs)
        assert max(s["seek"] for s in segs) > 0                                      # a second window was decoded
        assert out["text"] == decoding._text([t for s in segs for t in s["tokens"] if t < arch.eot])
        # with the default thresholds the fallback ladder runs (random-init text is "too repetitive"): still a result
        out2 = tr.transcribe(audio[0][:80000], m, language=None)
        assert isinstance(out2["text"], str) and out2["language"] in w.archs.LANGUAGES
    finally:
        decoding.set_detokenizer(None)
        m.close()
