"""GPU parity of the model path against the oracle (HF transformers on CPU + its pinned restatement):
fp32 path — encoder output, teacher-forced logits, and BIT-EXACT greedy ids (config 1: whisper-tiny, 16 clips, 220 tokens);
half-precision path (fp16 and bf16 builds) — logits within a stated relative error and token agreement reported."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w(built_lib):
    import whisper_ipa_b200 as w
    torch.backends.cuda.matmul.allow_tf32 = False
    return w


def _oracle_enc(sd, audio):
    from oracle import whisper_oracle as wo
    dims = wo.Dims.from_arch("tiny")
    mel = wo.log_mel_spectrogram(torch.from_numpy(audio), 80)
    return dims, mel, wo.encoder_forward(sd, dims, mel)


@pytest.mark.parametrize("gain", [1.0, 3.0])
def test_fp32_encoder_and_logits(w, tiny_sd, tiny_gain_sd, gain):
    from oracle import whisper_oracle as wo
    sd = tiny_sd if gain == 1.0 else tiny_gain_sd
    audio = wo.synthetic_audio(3)
    dims, mel_ref, enc_ref = _oracle_enc(sd, audio)
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=3)
    m.load_state_dict(sd)
    enc = m.encoder(w.log_mel_features(audio, 80)).cpu()
    assert (enc - enc_ref).abs().max().item() < 2e-3 * enc_ref.abs().max().item()
    # teacher-forced logits on the oracle's own encoder output isolate the decoder
    m.set_audio_features(enc_ref)
    toks = torch.tensor([wo.PROMPT_PRE_V3 + [1000 + 7 * i for i in range(6)]] * 3)
    ref = wo.decoder_forward(sd, dims, toks, 0, wo.cross_kv(sd, dims, enc_ref), [None] * dims.n_dec)
    got = m.teacher_forced_logits(toks).cpu()
    assert (got - ref).abs().max().item() < 1e-3 * ref.abs().max().item()
    assert torch.equal(got.argmax(-1), ref.argmax(-1))
    m.close()


def test_fp32_greedy_matches_golden(w, tiny_sd, tiny_gain_sd, golden_dir):
    from oracle import whisper_oracle as wo
    audio = wo.synthetic_audio(4)
    for name, sd in (("tiny_fp32", tiny_sd), ("tiny_gain_fp32", tiny_gain_sd)):
        g = np.load(os.path.join(golden_dir, f"{name}.npz"))
        m = w.WhisperIPA("tiny", dtype="float32", max_batch=4)
        m.load_state_dict(sd)
        ids = m.generate(w.log_mel_features(audio, 80), decoder_input_ids=torch.tensor([wo.PROMPT_PRE_V3] * 4),
                         max_new_tokens=24).cpu().numpy()
        assert (ids == g["greedy_ids"]).all(), name
        m.close()


def test_fp32_config1_bit_exact_ids_and_per(w, tiny_sd):
    """BASELINE config 1: whisper-tiny greedy decode of 16 synthetic 30 s clips (220 new tokens) + PER, vs HF generate."""
    from oracle import hf_reference as hf
    from oracle import per_oracle as po
    from oracle import whisper_oracle as wo
    from whisper_ipa_b200 import pipeline
    n = 16
    audio = wo.synthetic_audio(n)
    hf_model = hf.build_hf_model("tiny", seed=0)
    want = hf.hf_generate(hf_model, hf.hf_log_mel(audio, 80), "tiny", max_new=220)
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=8)        # two micro-batches
    m.load_state_dict(tiny_sd)
    refs = wo.synthetic_references(n)
    out = pipeline.Transcriber(m, max_new=220).evaluate_ids(audio, refs, micro_batch=8)
    hyps = out["local_hypotheses"]
    assert [len(h) for h in hyps] == [want.shape[1]] * n
    assert torch.equal(torch.tensor(hyps), want)
    want_d = po.levenshtein_batch(refs, [np.asarray(h, np.int32) for h in want.tolist()])
    assert (out["counts"][:, 0] == want_d).all()
    want_per = [po.per_from_counts(int(d), len(r), 220) for d, r in zip(want_d, refs)]
    assert out["per_scores"] == want_per and out["per"] == np.mean(want_per) and out["per_std"] == np.std(want_per)
    m.close()


@pytest.mark.parametrize("eot_like", [40220, 2020])
def test_eos_handling_matches_oracle(w, tiny_gain_sd, eot_like):
    """Rows that emit EOS stop at different steps, are padded with EOT and report their length; the batch stops once every
    row is done (HF:generation/utils.py:2796-2805).  EOS is made reachable by copying a frequent token's (tied) embedding
    into the EOT row; the HF-shaped generate() output must equal HF's own generate on the same weights."""
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    sd = {k: v.clone() for k, v in tiny_gain_sd.items()}
    emb = sd["model.decoder.embed_tokens.weight"]
    emb[wo.EOT] = emb[eot_like] * 1.05
    sd["proj_out.weight"] = emb
    dims = wo.Dims.from_arch("tiny")
    audio = wo.synthetic_audio(4)
    mel = wo.log_mel_spectrogram(torch.from_numpy(audio), 80)
    enc = wo.encoder_forward(sd, dims, mel)
    want, want_len = wo.greedy_decode(sd, dims, enc, wo.PROMPT_PRE_V3, 40)
    assert int(want_len.min()) < int(want_len.max()) < 40, "the crafted weights are meant to finish rows at different steps"
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=4)
    m.load_state_dict(sd)
    feats = w.log_mel_features(audio, 80)
    m.encoder(feats, return_features=False)
    ids, lens = m.decode_tokens(wo.PROMPT_PRE_V3, 40)
    assert torch.equal(lens.cpu().long(), want_len) and torch.equal(ids.cpu().long(), want)
    hf_model = hf.build_hf_model("tiny", seed=0, init_gain=3.0)
    with torch.no_grad():
        hf_model.model.decoder.embed_tokens.weight[wo.EOT] = emb[wo.EOT]
    hf_ids = hf.hf_generate(hf_model, hf.hf_log_mel(audio, 80), "tiny", max_new=40)
    mine = m.generate(feats, decoder_input_ids=torch.tensor([wo.PROMPT_PRE_V3] * 4), max_new_tokens=40).cpu()
    assert mine.shape == hf_ids.shape and torch.equal(mine, hf_ids)
    m.close()


# relative-L2 limits (encoder output, teacher-forced logits) per 16-bit type, <= 1.5 x the worst value measured on a B200
# over the six cases (round 2, gpurun_out/r2a_tests.log): fp16 encoder 2.08e-4 (tiny) / 3.17e-4 (base), logits 7.07e-4 - 7.23e-4
# (held to north_star's 1e-3); bf16 encoder 1.71e-3 / 2.61e-3, logits 5.86e-3 - 6.38e-3.  Token agreement was 1.000 everywhere.
HALF_LIMITS = {"float16": (4.8e-4, 1.0e-3), "bfloat16": (3.9e-3, 9.6e-3)}


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("arch,B,persistent,latent", [("tiny", 4, False, False), ("base", 2, False, False), ("tiny", 3, True, False),
                                                        ("base", 2, True, False), ("tiny", 4, False, True), ("base", 3, False, True)])
def test_half_logits_and_token_agreement(w, arch, B, persistent, latent, dtype, monkeypatch):
    """Half-precision path (fp16 = libwipa.so, bf16 = libwipa_bf16.so) vs the fp32 oracle: relative L2 error of the encoder
    output and of teacher-forced logits (north_star: 1e-3, met by fp16; bf16's 8-bit significand bounds it at the 6e-3
    level) and greedy-token agreement over 32 steps.  `persistent` forces every
    encoder GEMM through the persistent 128x256 kernel (normally chosen only for >= 2 waves of tiles) so that its
    fused epilogues (head split, residual add, fast GELU, conv rows) are checked at test sizes too."""
    if persistent:
        monkeypatch.setenv("WIPA_PERSISTENT_MIN_TILES", "1")
    # latent: cross-attention over the encoder output itself with the k / v projections folded (attn_lat.cu, the default);
    # otherwise the per-layer cross-KV cache and its stream-K kernel
    monkeypatch.setenv("WIPA_XATTN_LATENT", "1" if latent else "0")
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    sd = hf.state_dict_f32(hf.build_hf_model(arch, seed=0))
    dims = wo.Dims.from_arch(arch)
    audio = wo.synthetic_audio(B)
    mel_ref = wo.log_mel_spectrogram(torch.from_numpy(audio), 80)
    enc_ref = wo.encoder_forward(sd, dims, mel_ref)
    m = w.WhisperIPA(arch, dtype=dtype, max_batch=B)
    m.load_state_dict(sd)
    enc = m.encoder(w.log_mel_features(audio, 80)).cpu()
    enc_rel = ((enc - enc_ref).norm() / enc_ref.norm()).item()
    prompt = wo.PROMPT_PRE_V3
    ref_ids, _ = wo.greedy_decode(sd, dims, enc_ref, prompt, 32)
    toks = torch.cat([torch.tensor([prompt] * B), ref_ids[:, :-1]], dim=1)
    ref_logits = wo.decoder_forward(sd, dims, toks, 0, wo.cross_kv(sd, dims, enc_ref), [None] * dims.n_dec)
    got_logits = m.teacher_forced_logits(toks).cpu()
    rel = ((got_logits - ref_logits).norm() / ref_logits.norm()).item()
    ids, _ = m.decode_tokens(prompt, 32)
    agree = (ids.cpu().long() == ref_ids).float().mean().item()
    prefix = np.mean([int((row != ref).nonzero()[0]) if (row != ref).any() else 32
                      for row, ref in zip(ids.cpu().long(), ref_ids)])
    print(f"\n[{dtype} {arch}{' latent' if latent else ''}{' persistent' if persistent else ''}] encoder rel-L2 {enc_rel:.2e}, logits rel-L2 {rel:.2e}, token agreement {agree:.3f}, "
          f"mean agreeing prefix {prefix:.1f}/32")
    lim_enc, lim_logits = HALF_LIMITS[dtype]
    assert enc_rel < lim_enc and rel < lim_logits
    assert agree >= 0.97
    assert np.isfinite(got_logits.numpy()).all()
    m.close()


@pytest.mark.parametrize("beams", [2, 5])
@pytest.mark.parametrize("gain,eot_like,max_new", [(1.0, None, 24), (3.0, None, 20), (3.0, 40220, 40), (3.0, 2020, 40)])
def test_beam_search_matches_hf(w, tiny_sd, tiny_gain_sd, beams, gain, eot_like, max_new):
    """Beam search (HF `_beam_search` semantics: length_penalty 1.0, early_stopping False) on the fp32 path must return
    HF's own hypotheses: log-softmax + running scores, top-2K continuations, finished-slot bookkeeping with the length
    penalty, the early-stop heuristic, and beam reordering (done here through the ancestry table, never moving KV)."""
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    sd = {k: v.clone() for k, v in (tiny_sd if gain == 1.0 else tiny_gain_sd).items()}
    hf_model = hf.build_hf_model("tiny", seed=0, init_gain=gain)
    if eot_like is not None:                  # make EOS reachable (see test_eos_handling_matches_oracle)
        emb = sd["model.decoder.embed_tokens.weight"]
        emb[wo.EOT] = emb[eot_like] * 1.05
        sd["proj_out.weight"] = emb
        with torch.no_grad():
            hf_model.model.decoder.embed_tokens.weight[wo.EOT] = emb[wo.EOT]
    B = 3
    audio = wo.synthetic_audio(B)
    want = hf.hf_generate(hf_model, hf.hf_log_mel(audio, 80), "tiny", max_new=max_new, num_beams=beams)
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=B, max_beams=beams)
    m.load_state_dict(sd)
    got = m.generate(w.log_mel_features(audio, 80), decoder_input_ids=torch.tensor([wo.PROMPT_PRE_V3] * B),
                     max_new_tokens=max_new, num_beams=beams, length_penalty=1.0).cpu()
    m.close()
    width = max(got.shape[1], want.shape[1])
    pad = lambda t: torch.nn.functional.pad(t, (0, width - t.shape[1]), value=wo.EOT)
    assert torch.equal(pad(got), pad(want)), f"beam search differs from HF:\n{got}\n{want}"
    if eot_like is not None:
        assert want.shape[1] < max_new, "the crafted weights are meant to finish hypotheses before max_new"


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("latent", [False, True])
def test_beam_search_half_runs_and_mostly_agrees(w, tiny_gain_sd, latent, dtype, monkeypatch):
    """bf16 path through the same beam bookkeeping: the logits differ from the fp32 oracle at the 1e-2 level, so the
    hypotheses are compared as token agreement (reported), not exactly.  `latent`: cross-attention over the encoder
    output (the beams of an utterance share it through utt_of_seq) instead of the per-layer cross-KV."""
    monkeypatch.setenv("WIPA_XATTN_LATENT", "1" if latent else "0")
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    B, beams, max_new = 4, 5, 24
    audio = wo.synthetic_audio(B)
    hf_model = hf.build_hf_model("tiny", seed=0, init_gain=3.0)
    want = hf.hf_generate(hf_model, hf.hf_log_mel(audio, 80), "tiny", max_new=max_new, num_beams=beams)
    m = w.WhisperIPA("tiny", dtype=dtype, max_batch=B, max_beams=beams)
    m.load_state_dict(tiny_gain_sd)
    got = m.generate(w.log_mel_features(audio, 80), decoder_input_ids=torch.tensor([wo.PROMPT_PRE_V3] * B),
                     max_new_tokens=max_new, num_beams=beams).cpu()
    m.close()
    assert got.shape[0] == B and got.shape[1] <= max_new
    n = min(got.shape[1], want.shape[1])
    agree = (got[:, :n] == want[:, :n]).float().mean().item()
    print(f"\n[{dtype} beam {beams}{' latent' if latent else ''}] token agreement with HF fp32 beam search {agree:.3f}")
    assert agree > 0.8


@pytest.mark.parametrize("shape", ["large-v3", "medium"])
def test_wide_architectures_reduced_depth(w, shape):
    """Shape generality: the large-v3 geometry (d 1280, 20 heads, ffn 5120, 128 mel bins, vocab 51866, +1-shifted prompt
    ids) and the medium geometry (d 1024, 16 heads), cut to 2+2 layers so the HF oracle stays cheap.  fp32 greedy ids
    must equal HF's; the bf16 path reports token agreement."""
    import logging
    import warnings
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from whisper_ipa_b200.archs import ARCHS, WhisperArch
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    full = ARCHS[shape]
    arch = WhisperArch(full.name, full.d_model, 2, 2, full.heads, full.ffn, full.n_mels, full.vocab)
    logging.getLogger("transformers").setLevel(logging.ERROR)
    cfg = WhisperConfig(**arch.hf_config_kwargs(), decoder_start_token_id=50258, eos_token_id=50257, pad_token_id=50257,
                        bos_token_id=50257, begin_suppress_tokens=[220, 50257])
    torch.manual_seed(0)
    hf_model = WhisperForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        for n, p in hf_model.named_parameters():
            if p.dim() >= 2 and "embed_positions" not in n:
                p.mul_(2.0)                               # make the ids depend on the audio (see build_hf_model)
    sd = hf.state_dict_f32(hf_model)
    B, max_new = 2, 10
    audio = wo.synthetic_audio(B)
    prompt = arch.prompt("en", "transcribe", True)
    feats_hf = hf.hf_log_mel(audio, arch.n_mels)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = hf_model.generate(feats_hf, decoder_input_ids=torch.tensor([prompt] * B), max_new_tokens=max_new, do_sample=False)
    feats = w.log_mel_features(audio, arch.n_mels)
    assert (feats.cpu() - feats_hf).abs().max().item() < 1e-3
    for dtype in ("float32", "float16", "bfloat16"):
        m = w.WhisperIPA(arch, dtype=dtype, max_batch=B)
        m.load_state_dict(sd)
        got = m.generate(feats, decoder_input_ids=torch.tensor([prompt] * B), max_new_tokens=max_new).cpu()
        m.close()
        n = min(got.shape[1], want.shape[1])
        agree = (got[:, :n] == want[:, :n]).float().mean().item()
        print(f"\n[{shape} geometry, 2+2 layers, {dtype}] token agreement with HF {agree:.3f}")
        if dtype == "float32":
            assert got.shape == want.shape and torch.equal(got, want)
        else:
            assert agree > 0.7


def test_fp32_maximum_target_length(w, tiny_gain_sd):
    """Decode to the architectural limit (4 prompt + 444 new = 448 positions = all 28 pages of the self-KV block
    table): ids must still equal the oracle's, position by position."""
    from oracle import whisper_oracle as wo
    dims = wo.Dims.from_arch("tiny")
    audio = wo.synthetic_audio(1)
    mel = wo.log_mel_spectrogram(torch.from_numpy(audio), 80)
    enc = wo.encoder_forward(tiny_gain_sd, dims, mel)
    max_new = 448 - len(wo.PROMPT_PRE_V3)
    want, _ = wo.greedy_decode(tiny_gain_sd, dims, enc, wo.PROMPT_PRE_V3, max_new)
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=1)
    m.load_state_dict(tiny_gain_sd)
    m.set_audio_features(enc)
    ids, lens = m.decode_tokens(wo.PROMPT_PRE_V3, max_new)
    m.close()
    assert ids.shape == (1, max_new) and torch.equal(ids.cpu().long(), want)
    with pytest.raises(Exception):
        m2 = w.WhisperIPA("tiny", dtype="float32", max_batch=1)
        try:
            m2.load_state_dict(tiny_gain_sd)
            m2.set_audio_features(enc)
            m2.decode_tokens(wo.PROMPT_PRE_V3, max_new + 1)          # 449 positions: refused, not truncated
        finally:
            m2.close()


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_half_decode_is_deterministic_and_batch_stable(w, tiny_gain_sd, dtype):
    """Two runs over the same 37 clips give identical ids (the stream-K cross-attention merges partials in a fixed
    order), and the ids of a clip do not depend on which other clips share its batch (a ragged last micro-batch)."""
    from oracle import whisper_oracle as wo
    B = 37
    audio = wo.synthetic_audio(B)
    feats = w.log_mel_features(audio, 80)
    m = w.WhisperIPA("tiny", dtype=dtype, max_batch=B)
    m.load_state_dict(tiny_gain_sd)
    prompt = torch.tensor([wo.PROMPT_PRE_V3] * B)
    a = m.generate(feats, decoder_input_ids=prompt, max_new_tokens=24).cpu()
    b = m.generate(feats, decoder_input_ids=prompt, max_new_tokens=24).cpu()
    assert torch.equal(a, b)
    m.close()
    m = w.WhisperIPA("tiny", dtype=dtype, max_batch=16)                  # micro-batches of 16, 16, 5
    m.load_state_dict(tiny_gain_sd)
    c = m.generate(feats, decoder_input_ids=prompt, max_new_tokens=24).cpu()
    m.close()
    agree = (a == c).float().mean().item()
    print(f"\n[{dtype} batch stability] token agreement between batch 37 and micro-batches of 16: {agree:.4f}")
    assert agree > 0.98
