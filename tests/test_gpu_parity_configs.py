"""Parity of the half-precision production path ON THE CONFIGURATIONS THE BENCHMARK RUNS (BASELINE configs 2 and 3):
whisper-small with >= 128 sequences and whisper-base with 64, no environment overrides - i.e. the defaults the bench uses
(latent cross-attention from 96 sequences on, persistent tcgen05 GEMMs, tcgen05 flash attention, CUDA-graph decode).

Oracle: HF transformers' own fp32 path run on the same B200 with TF32 disabled (oracle/hf_reference.hf_gpu_fp32_reference:
stock WhisperFeatureExtractor, encoder, generate() and a teacher-forced decoder pass).

north_star: "mel features and logits within 1e-3 relative error in the [half-precision] path, with token-sequence
agreement reported".  Every assertion below is <= 1.5 x the value measured on a B200 (stated next to it); the fp16 build
(libwipa.so, the default) must meet 1e-3, the bf16 build (libwipa_bf16.so) is held to its own measured level.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# (arch, clips, dtype) -> (mel max-abs, encoder rel-L2, logits rel-L2, teacher-forced argmax agreement >=, free-running
# token agreement >=).  Measured on a B200 (round 2, final build: folded LayerNorm, two-step latent projections):
#   small B=128 fp16: mel 8.1e-5, encoder 4.98e-4, logits 8.41e-4 (worst row 8.58e-4), agreement 1.0 / 1.0, 128/128 rows identical
#   small B=128 bf16: encoder 4.10e-3, logits 6.84e-3, agreement 1.0 / 1.0
#   base  B=64  fp16: encoder 3.17e-4, logits 7.03e-4, agreement 1.0 / 1.0;   bf16: encoder 2.61e-3, logits 5.78e-3
# Limits are 1.5 x measured, except fp16 logits, which are held to north_star's 1e-3 itself (1.18 x / 1.41 x measured).
LIMITS = {
    ("small", 128, "float16"): (1.3e-4, 7.5e-4, 1.0e-3, 0.99, 0.97),
    ("small", 128, "bfloat16"): (1.3e-4, 6.2e-3, 1.0e-2, 0.99, 0.97),
    ("base", 64, "float16"): (1.3e-4, 4.8e-4, 1.0e-3, 0.99, 0.97),
    ("base", 64, "bfloat16"): (1.3e-4, 3.9e-3, 8.8e-3, 0.99, 0.97),
}
MAX_NEW = 220
LOGIT_STEPS = 48
LOGIT_ROWS = 32


@pytest.fixture(scope="module")
def w(built_lib):
    import whisper_ipa_b200 as w
    return w


_ref_cache = {}


def _reference(arch, n):
    key = (arch, n)
    if key not in _ref_cache:
        from oracle import hf_reference as hf
        from oracle import whisper_oracle as wo
        audio = wo.synthetic_audio(n)
        hf_model = hf.build_hf_model(arch, seed=0)
        ref = hf.hf_gpu_fp32_reference(hf_model, audio, wo.prompt_for(arch), MAX_NEW, LOGIT_STEPS, logit_rows=LOGIT_ROWS)
        ref["audio"] = audio
        ref["sd"] = hf.state_dict_f32(hf_model)
        _ref_cache.clear()                       # one architecture resident at a time
        _ref_cache[key] = ref
    return _ref_cache[key]


@pytest.mark.parametrize("arch,n,dtype", list(LIMITS))
def test_half_precision_path_on_bench_config(w, arch, n, dtype, monkeypatch):
    for var in ("WIPA_XATTN_LATENT", "WIPA_PERSISTENT_MIN_TILES", "WIPA_BN_DEC", "WIPA_NO_GRAPH", "WIPA_ENC_ATTN_SIMT"):
        monkeypatch.delenv(var, raising=False)
    from oracle import whisper_oracle as wo
    ref = _reference(arch, n)
    lim_mel, lim_enc, lim_logits, lim_agree, lim_free = LIMITS[(arch, n, dtype)]
    prompt = wo.prompt_for(arch)

    m = w.WhisperIPA(arch, dtype=dtype, max_batch=n)
    m.load_state_dict(ref["sd"])
    assert m.info()["xattn_latent"] == (1 if n >= 96 else 0), "the default cross-attention selection changed"
    mel = w.log_mel_features(ref["audio"], m.arch.n_mels)
    mel_err = (mel.cpu() - ref["mel"]).abs().max().item()
    enc = m.encoder(mel).cpu()
    enc_rel = ((enc - ref["enc"]).norm() / ref["enc"].norm()).item()

    # teacher-forced logits on the oracle's own greedy tokens (first LOGIT_ROWS utterances are compared; the library
    # computes all n rows, so the decode GEMMs run at the benchmark's M)
    got = m.teacher_forced_logits(ref["tokens"])[:LOGIT_ROWS].cpu()
    want = ref["logits"]
    rel = ((got - want).norm() / want.norm()).item()
    rel_worst_row = ((got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max().item()
    P = len(prompt)
    tf_agree = (got[:, P - 1:].argmax(-1) == want[:, P - 1:].argmax(-1)).float().mean().item()

    # free-running greedy decode through the CUDA graph, all n clips, 220 tokens
    m.encoder(mel, return_features=False)
    ids, lens = m.decode_tokens(prompt, MAX_NEW)
    ids = ids.cpu().long()
    want_ids = ref["ids"]
    L = min(ids.shape[1], want_ids.shape[1])
    agree = (ids[:, :L] == want_ids[:, :L]).float().mean().item()
    neq = ids[:, :L] != want_ids[:, :L]
    prefix = np.mean([int(r.nonzero()[0]) if r.any() else L for r in neq])
    exact_rows = int((~neq.any(dim=1)).sum())
    m.close()
    print(f"\n[parity {arch} B={n} {dtype}] mel max-abs {mel_err:.2e}; encoder rel-L2 {enc_rel:.2e}; logits rel-L2 {rel:.2e} "
          f"(worst row {rel_worst_row:.2e}); teacher-forced argmax agreement {tf_agree:.4f} over {LOGIT_ROWS}x{LOGIT_STEPS}; "
          f"free-running: token agreement {agree:.4f}, mean agreeing prefix {prefix:.1f}/{L}, {exact_rows}/{n} rows identical")
    assert np.isfinite(got.numpy()).all()
    assert mel_err < lim_mel
    assert enc_rel < lim_enc
    assert rel < lim_logits
    assert tf_agree >= lim_agree
    assert agree >= lim_free
    assert lens.cpu().tolist() == [want_ids.shape[1]] * n or want_ids.shape[1] < MAX_NEW


# ---- BASELINE configs 4 and 5 at FULL depth ---------------------------------------------------------------------------------
def _hf_gpu_generate(hf_model, audio, prompt, max_new, n_mels, **kw):
    """HF fp32 generate on the GPU (TF32 off): ids int64 [B, <= max_new] on the CPU."""
    import warnings
    from oracle import hf_reference as hf
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        feats = hf.hf_log_mel(audio, n_mels).cuda()
        m = hf_model.cuda().float().eval()
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ids = m.generate(feats, decoder_input_ids=torch.tensor([list(prompt)] * len(audio), device="cuda"), max_new_tokens=max_new,
                             do_sample=False, **kw)
        return ids.cpu()
    finally:
        hf_model.cpu()
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_medium_beam5_full_depth_matches_hf(w):
    """BASELINE configs[3]: whisper-medium (24 + 24 layers, 16 heads), beam search with 5 beams, length_penalty 1.0.
    fp32 path: the hypotheses must be HF's own (`_beam_search`); fp16 path: token agreement reported."""
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    arch = w.ARCHS["medium"]
    hf_model = hf.build_hf_model("medium", seed=0, init_gain=2.0)          # gain: hypotheses depend on the audio
    sd = hf.state_dict_f32(hf_model)
    B, beams, max_new = 2, 5, 16
    audio = wo.synthetic_audio(B)
    prompt = arch.prompt("en", "transcribe", True)
    want = _hf_gpu_generate(hf_model, audio, prompt, max_new, 80, num_beams=beams, length_penalty=1.0, early_stopping=False)
    feats = w.log_mel_features(audio, 80)
    for dtype in ("float32", "float16"):
        m = w.WhisperIPA("medium", dtype=dtype, max_batch=B, max_beams=beams)
        m.load_state_dict(sd)
        got = m.generate(feats, decoder_input_ids=torch.tensor([prompt] * B), max_new_tokens=max_new, num_beams=beams,
                         length_penalty=1.0).cpu()
        m.close()
        width = max(got.shape[1], want.shape[1])
        pad = lambda t: torch.nn.functional.pad(t, (0, width - t.shape[1]), value=arch.eot)
        agree = (pad(got) == pad(want)).float().mean().item()
        print(f"\n[medium full depth, beam {beams}, {dtype}] token agreement with HF fp32 beam search {agree:.3f}")
        if dtype == "float32":
            assert torch.equal(pad(got), pad(want)), f"beam search differs from HF:\n{got}\n{want}"
        else:
            # measured on a B200: 0.719 - one of the two utterances keeps HF's hypothesis, the other one's best beam switches at
            # a near-tie of cumulative scores (random-init logits are almost flat) and diverges from there; reported, loosely bound
            assert agree >= 0.5


def test_large_v3_full_depth_greedy_and_per_match_hf(w, monkeypatch):
    """BASELINE configs[4]: whisper-large-v3 (32 + 32 layers, 20 heads, 128 mel bins, vocab 51866): greedy ids of the fp32
    path equal HF's, PER counts equal the CPU oracle's; the fp16 path reports token agreement twice - with the stream-K
    cross-attention over per-layer K/V (what a 2-clip context picks) and with the latent cross-attention forced (what contexts
    of 96 sequences and more pick: csrc/attn_lat_wide.cu, two CTAs of 10 heads per key range)."""
    from oracle import hf_reference as hf
    from oracle import per_oracle as po
    from oracle import whisper_oracle as wo
    from whisper_ipa_b200 import metrics
    arch = w.ARCHS["large-v3"]
    hf_model = hf.build_hf_model("large-v3", seed=0, init_gain=1.5)
    sd = hf.state_dict_f32(hf_model)
    B, max_new = 2, 16
    audio = wo.synthetic_audio(B)
    prompt = arch.prompt("en", "transcribe", True)
    assert prompt == wo.PROMPT_V3
    want = _hf_gpu_generate(hf_model, audio, prompt, max_new, 128)
    feats = w.log_mel_features(audio, 128)
    refs = wo.synthetic_references(B)
    for dtype, latent in (("float32", None), ("float16", "0"), ("float16", "1")):
        if latent is not None:
            monkeypatch.setenv("WIPA_XATTN_LATENT", latent)
        m = w.WhisperIPA("large-v3", dtype=dtype, max_batch=B)
        if latent is not None:
            assert m.info()["xattn_latent"] == int(latent)
        m.load_state_dict(sd)
        got = m.generate(feats, decoder_input_ids=torch.tensor([prompt] * B), max_new_tokens=max_new).cpu()
        m.close()
        n = min(got.shape[1], want.shape[1])
        agree = (got[:, :n] == want[:, :n]).float().mean().item()
        print(f"\n[large-v3 full depth, greedy, {dtype}, latent={latent}] token agreement with HF fp32 {agree:.3f}")
        if dtype == "float32":
            assert got.shape == want.shape and torch.equal(got, want)
            counts = metrics.edit_distance_counts(refs, [r.tolist() for r in got]).cpu().numpy()
            assert (counts[:, 0] == po.levenshtein_batch(refs, [np.asarray(r, np.int32) for r in want.tolist()])).all()
        else:
            assert agree >= 0.9
