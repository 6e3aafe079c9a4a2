"""CPU restatement of the Whisper hot path (log-mel -> encoder -> KV-cached greedy / beam decode).

TEST INFRASTRUCTURE ONLY.  Nothing under ``whisper_ipa_b200/`` may import this module; it is
imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the *checker*, never as the product.

What it restates
----------------
The reference scripts (``ref:scripts/evaluate_model.py:187-201``, ``ref:scripts/transcribe_single.py:43-56``)
delegate every arithmetic step to ``mlx-whisper==0.4.3`` (``ref:requirements.txt:23``), which is NOT vendored
under /root/reference and cannot be installed here (needs Apple Metal).  BASELINE.json names the
HF ``WhisperForConditionalGeneration.generate`` path as the parity oracle, so this file restates
*that* published algorithm in plain fp32 torch-CPU tensor code, function by function:

* ``mel_filter_bank``      <- HF:audio_utils.py:453-544 (slaney scale, slaney norm)
* ``log_mel_spectrogram``  <- HF:models/whisper/feature_extraction_whisper.py:135-164
* ``sinusoids``            <- HF:models/whisper/modeling_whisper.py:55-65
* ``encoder_forward``      <- HF:models/whisper/modeling_whisper.py:593-647 (+ layer :361-414, attention :284-357)
* ``decoder_step``         <- HF:models/whisper/modeling_whisper.py:449-506, 691-796, 1081
* ``greedy_decode``        <- HF:generation/utils.py:2658-2841 with the begin-suppress processor
                              HF:generation/logits_process.py:1855-1902
* ``beam_decode``          <- HF:generation/utils.py:3076-3400 (length_penalty 1.0, early_stopping False)

Pinning: ``tests/test_oracle_whisper.py`` checks every function here against the real HF
implementation (importable in this image) and against ``tests/golden/*.npz`` fixtures generated
from HF by ``oracle/make_golden.py``.  Parity is therefore pinned to HF 5.5.0; the reference tree
itself holds no golden tensor for this path ("parity unpinned" on the reference side).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_AUDIO_CTX = 1500

ARCHS: Dict[str, Dict[str, int]] = {
    # HF WhisperConfig kwargs for the five named architectures
    "tiny": dict(d_model=384, encoder_layers=4, decoder_layers=4, encoder_attention_heads=6,
                 decoder_attention_heads=6, encoder_ffn_dim=1536, decoder_ffn_dim=1536,
                 num_mel_bins=80, vocab_size=51865),
    "base": dict(d_model=512, encoder_layers=6, decoder_layers=6, encoder_attention_heads=8,
                 decoder_attention_heads=8, encoder_ffn_dim=2048, decoder_ffn_dim=2048,
                 num_mel_bins=80, vocab_size=51865),
    "small": dict(d_model=768, encoder_layers=12, decoder_layers=12, encoder_attention_heads=12,
                  decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072,
                  num_mel_bins=80, vocab_size=51865),
    "medium": dict(d_model=1024, encoder_layers=24, decoder_layers=24, encoder_attention_heads=16,
                   decoder_attention_heads=16, encoder_ffn_dim=4096, decoder_ffn_dim=4096,
                   num_mel_bins=80, vocab_size=51865),
    "large-v3": dict(d_model=1280, encoder_layers=32, decoder_layers=32, encoder_attention_heads=20,
                     decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120,
                     num_mel_bins=128, vocab_size=51866),
}

PROMPT_PRE_V3 = [50258, 50259, 50359, 50363]   # <|sot|><|en|><|transcribe|><|notimestamps|>
PROMPT_V3 = [50258, 50259, 50360, 50364]
EOT = 50257
BEGIN_SUPPRESS = [220, 50257]


def prompt_for(arch: str) -> List[int]:
    return PROMPT_V3 if arch == "large-v3" else PROMPT_PRE_V3


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_audio(n_clips: int, first: int = 0) -> np.ndarray:
    """30 s clips of 0.1-scaled Gaussian noise, seed 1234 + clip index."""
    out = np.empty((n_clips, N_SAMPLES), dtype=np.float32)
    for i in range(n_clips):
        out[i] = np.random.default_rng(1234 + first + i).standard_normal(N_SAMPLES).astype(np.float32) * 0.1
    return out


def synthetic_references(n_clips: int, first: int = 0) -> List[np.ndarray]:
    """PER reference id sequences: one rng(4321) stream, consumed in clip order."""
    rng = np.random.default_rng(4321)
    seqs = []
    for _ in range(first + n_clips):
        n = int(rng.integers(20, 120))
        seqs.append(rng.integers(0, 50257, size=n).astype(np.int32))
    return seqs[first:]


# --------------------------------------------------------------------------------------
# log-mel front end
# --------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    big = f >= 1000.0
    mels = np.where(big, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    big = m >= 15.0
    return np.where(big, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def mel_filter_bank(n_mels: int, n_freq: int = 201, sr: int = 16000, fmax: float = 8000.0) -> np.ndarray:
    """[n_freq, n_mels] float64 triangular slaney filterbank (HF:audio_utils.py:453-544)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def log_mel_spectrogram(audio: torch.Tensor, n_mels: int = 80) -> torch.Tensor:
    """audio f32[B, 480000] -> f32[B, n_mels, 3000]  (HF:...feature_extraction_whisper.py:135-164)."""
    audio = torch.as_tensor(audio, dtype=torch.float32)
    if audio.dim() == 1:
        audio = audio[None]
    B = audio.shape[0]
    window = torch.hann_window(N_FFT, periodic=True, dtype=torch.float32)
    padded = F.pad(audio[:, None, :], (N_FFT // 2, N_FFT // 2), mode="reflect")[:, 0]
    frames = padded.unfold(-1, N_FFT, HOP)                      # [B, 3001, 400]
    spec = torch.fft.rfft(frames * window, dim=-1)              # [B, 3001, 201]
    power = (spec.real ** 2 + spec.imag ** 2)[:, :-1, :]         # drop last frame -> [B, 3000, 201]
    fb = torch.from_numpy(mel_filter_bank(n_mels)).to(torch.float32)    # [201, n_mels]
    mel = torch.matmul(power, fb).transpose(1, 2)               # [B, n_mels, 3000]
    log_spec = torch.clamp(mel, min=1e-10).log10()
    mx = log_spec.reshape(B, -1).max(dim=1).values.view(B, 1, 1)
    log_spec = torch.maximum(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


# --------------------------------------------------------------------------------------
# model restatement (functional, keyed by the HF state_dict names)
# --------------------------------------------------------------------------------------
def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


@dataclass
class Dims:
    d: int
    n_enc: int
    n_dec: int
    heads: int
    ffn: int
    n_mels: int
    vocab: int

    @staticmethod
    def from_arch(arch: str) -> "Dims":
        a = ARCHS[arch]
        return Dims(a["d_model"], a["encoder_layers"], a["decoder_layers"], a["encoder_attention_heads"],
                    a["encoder_ffn_dim"], a["num_mel_bins"], a["vocab_size"])


def _lin(x, sd, name, bias=True):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"] if bias else None)


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _heads(x, H):
    B, T, D = x.shape
    return x.view(B, T, H, D // H).transpose(1, 2)         # [B,H,T,hd]


def _attend(q, k, v, causal_offset: Optional[int] = None):
    """softmax(q k^T) v with NO extra scaling (q is pre-scaled, HF:...modeling_whisper.py:310,349)."""
    s = torch.matmul(q, k.transpose(-1, -2))
    if causal_offset is not None:
        Tq, Tk = s.shape[-2:]
        i = torch.arange(Tq).view(-1, 1) + causal_offset
        j = torch.arange(Tk).view(1, -1)
        s = s.masked_fill(j > i, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v)                                   # [B,H,Tq,hd]
    B, H, Tq, hd = o.shape
    return o.transpose(1, 2).reshape(B, Tq, H * hd)


def encoder_forward(sd: Dict[str, torch.Tensor], dims: Dims, mel: torch.Tensor,
                    taps: Optional[dict] = None) -> torch.Tensor:
    """mel f32[B, n_mels, 3000] -> f32[B, 1500, d]."""
    p = "model.encoder."
    x = F.gelu(F.conv1d(mel, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + sd[p + "embed_positions.weight"]
    if taps is not None:
        taps["stem"] = x.clone()
    scale = (dims.d // dims.heads) ** -0.5
    for l in range(dims.n_enc):
        lp = f"{p}layers.{l}."
        h = _ln(x, sd, lp + "self_attn_layer_norm")
        q = _heads(_lin(h, sd, lp + "self_attn.q_proj") * scale, dims.heads)
        k = _heads(_lin(h, sd, lp + "self_attn.k_proj", bias=False), dims.heads)
        v = _heads(_lin(h, sd, lp + "self_attn.v_proj"), dims.heads)
        x = x + _lin(_attend(q, k, v), sd, lp + "self_attn.out_proj")
        h = _ln(x, sd, lp + "final_layer_norm")
        x = x + _lin(F.gelu(_lin(h, sd, lp + "fc1")), sd, lp + "fc2")
        if taps is not None:
            taps[f"layer{l}"] = x.clone()
    return _ln(x, sd, p + "layer_norm")


def cross_kv(sd, dims: Dims, enc: torch.Tensor) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Per decoder layer (K, V) as [B,H,1500,hd] (HF:...modeling_whisper.py:326-336)."""
    out = []
    for l in range(dims.n_dec):
        lp = f"model.decoder.layers.{l}.encoder_attn."
        k = _heads(_lin(enc, sd, lp + "k_proj", bias=False), dims.heads)
        v = _heads(_lin(enc, sd, lp + "v_proj"), dims.heads)
        out.append((k, v))
    return out


def fold_cross_attention(sd, dims: Dims, layer: int):
    """Folded cross-attention projections of one decoder layer, the algebra libwipa's latent cross-attention kernel
    (whisper_ipa_b200/csrc/attn_lat.cu, weights built by ctx.cu:xlat_fold_*_kernel) relies on.  With E the encoder output,
    HF computes K = E Wk^T (no bias), V = E Wv^T + bv, q = (h Wq^T + bq) * scale, ctx_h = softmax(q_h K_h^T) V_h and
    out = ctx Wo^T + bo (HF:models/whisper/modeling_whisper.py:241-357).  K and V are linear in E, hence
        q_h . K_h[t]  = (Wk_h^T q_h) . E[t]                       -> q' = h Wq'^T + bq',   Wq'[(h,n),k] = s * sum_j Wk[64h+j,n] Wq[64h+j,k]
        ctx_h         = Wv_h (sum_t p_h[t] E[t]) + bv_h           -> out = C Wo'^T + bo',  Wo'[m,(h,n)] = sum_j Wo[m,64h+j] Wv[64h+j,n]
    with C[h] = sum_t p_h[t] E[t] and bo' = bo + Wo bv (sum_t p = 1).  Returns (Wq' [H*d, d], bq' [H*d], Wo' [d, H*d], bo' [d])."""
    lp = f"model.decoder.layers.{layer}.encoder_attn."
    d, H = dims.d, dims.heads
    hd = d // H
    scale = hd ** -0.5
    Wq, bq = sd[lp + "q_proj.weight"], sd[lp + "q_proj.bias"]
    Wk = sd[lp + "k_proj.weight"]
    Wv, bv = sd[lp + "v_proj.weight"], sd[lp + "v_proj.bias"]
    Wo, bo = sd[lp + "out_proj.weight"], sd[lp + "out_proj.bias"]
    Wq_h, Wk_h, Wv_h = Wq.view(H, hd, d), Wk.view(H, hd, d), Wv.view(H, hd, d)
    Wq2 = scale * torch.einsum("hjn,hjk->hnk", Wk_h, Wq_h).reshape(H * d, d)
    bq2 = scale * torch.einsum("hjn,hj->hn", Wk_h, bq.view(H, hd)).reshape(H * d)
    Wo2 = torch.einsum("mhj,hjn->mhn", Wo.view(d, H, hd), Wv_h).reshape(d, H * d)
    bo2 = bo + Wo @ bv
    return Wq2, bq2, Wo2, bo2


def cross_attention_latent(sd, dims: Dims, layer: int, h: torch.Tensor, enc: torch.Tensor) -> torch.Tensor:
    """Cross-attention block output (before the residual add) computed over the encoder output itself:
    h f32[B,T,d] (the LayerNorm output), enc f32[B,1500,d] -> f32[B,T,d].  Must equal
    out_proj(_attend(q_proj(h) * scale, k_proj(enc), v_proj(enc))) up to rounding."""
    Wq2, bq2, Wo2, bo2 = fold_cross_attention(sd, dims, layer)
    B, T, d = h.shape
    H = dims.heads
    qp = F.linear(h, Wq2, bq2).view(B, T, H, d)                       # absorbed queries, one R^d vector per head
    s = torch.einsum("bthd,bkd->bhtk", qp, enc)
    p = torch.softmax(s, dim=-1)
    c = torch.einsum("bhtk,bkd->bthd", p, enc).reshape(B, T, H * d)   # per-head averages of the encoder output
    return F.linear(c, Wo2, bo2)


def decoder_forward(sd, dims: Dims, tokens: torch.Tensor, pos0: int, xkv, self_kv: list) -> torch.Tensor:
    """tokens int64[B,T] at positions pos0.. -> logits f32[B,T,V]; appends to self_kv in place."""
    p = "model.decoder."
    B, T = tokens.shape
    x = sd[p + "embed_tokens.weight"][tokens] + sd[p + "embed_positions.weight"][pos0:pos0 + T]
    scale = (dims.d // dims.heads) ** -0.5
    for l in range(dims.n_dec):
        lp = f"{p}layers.{l}."
        h = _ln(x, sd, lp + "self_attn_layer_norm")
        q = _heads(_lin(h, sd, lp + "self_attn.q_proj") * scale, dims.heads)
        k = _heads(_lin(h, sd, lp + "self_attn.k_proj", bias=False), dims.heads)
        v = _heads(_lin(h, sd, lp + "self_attn.v_proj"), dims.heads)
        if self_kv[l] is None:
            self_kv[l] = (k, v)
        else:
            self_kv[l] = (torch.cat([self_kv[l][0], k], dim=2), torch.cat([self_kv[l][1], v], dim=2))
        x = x + _lin(_attend(q, self_kv[l][0], self_kv[l][1], causal_offset=pos0), sd, lp + "self_attn.out_proj")
        h = _ln(x, sd, lp + "encoder_attn_layer_norm")
        q = _heads(_lin(h, sd, lp + "encoder_attn.q_proj") * scale, dims.heads)
        x = x + _lin(_attend(q, xkv[l][0], xkv[l][1]), sd, lp + "encoder_attn.out_proj")
        h = _ln(x, sd, lp + "final_layer_norm")
        x = x + _lin(F.gelu(_lin(h, sd, lp + "fc1")), sd, lp + "fc2")
    x = _ln(x, sd, p + "layer_norm")
    return F.linear(x, sd[p + "embed_tokens.weight"])       # tied head, no bias (:966,1081)


def greedy_decode(sd, dims: Dims, enc: torch.Tensor, prompt: Sequence[int], max_new: int,
                  eot: int = EOT, begin_suppress: Sequence[int] = BEGIN_SUPPRESS,
                  suppress: Sequence[int] = (), return_logits: bool = False):
    """HF greedy semantics: returns int64[B, max_new] right-padded with ``eot`` (= pad id) after EOS,
    and the per-row count of tokens before EOS."""
    B = enc.shape[0]
    xkv = cross_kv(sd, dims, enc)
    self_kv = [None] * dims.n_dec
    toks = torch.tensor([list(prompt)] * B, dtype=torch.long)
    logits = decoder_forward(sd, dims, toks, 0, xkv, self_kv)[:, -1].float()
    out = torch.full((B, max_new), eot, dtype=torch.long)
    lens = torch.full((B,), max_new, dtype=torch.long)
    done = torch.zeros(B, dtype=torch.bool)
    all_logits = []
    pos = len(prompt)
    for i in range(max_new):
        logits = logits.clone()
        if len(suppress):
            logits[:, list(suppress)] = float("-inf")
        if i == 0 and len(begin_suppress):
            logits[:, list(begin_suppress)] = float("-inf")
        if return_logits:
            all_logits.append(logits.clone())
        nxt = logits.argmax(dim=-1)
        nxt = torch.where(done, torch.full_like(nxt, eot), nxt)
        newly = (~done) & (nxt == eot)
        lens = torch.where(newly, torch.full_like(lens, i), lens)
        done = done | newly
        out[:, i] = nxt
        if bool(done.all()) or i == max_new - 1:
            break
        logits = decoder_forward(sd, dims, nxt[:, None], pos, xkv, self_kv)[:, -1].float()
        pos += 1
    if return_logits:
        return out, lens, all_logits
    return out, lens


def beam_decode(sd, dims: Dims, enc: torch.Tensor, prompt: Sequence[int], max_new: int, beams: int,
                length_penalty: float = 1.0, eot: int = EOT,
                begin_suppress: Sequence[int] = BEGIN_SUPPRESS):
    """Beam search restating HF's vectorised `_beam_search` (HF:generation/utils.py:3076-3400 and its helpers
    :2876-3073) for do_sample=False, early_stopping=False, one EOS id, num_return_sequences=1:
      * log_softmax over the vocabulary, THEN the suppress processor (:3262-3263), plus the running beam scores (:3286)
      * top 2*beams continuations over beams x vocab (:2945-2997); a candidate "hits the stopping criteria" when its
        token is EOS or the sequence reaches max_length (:3304-3310)
      * next running beams = best `beams` of (score - 1e9 * hit) (:2999-3019)
      * finished slots = best `beams` of [old finished scores, candidate / (generated_len ** length_penalty) - 1e9 *
        (not among the top `beams` candidates or not hit or the utterance can no longer improve)] (:3021-3073)
      * early-stop heuristic (:2876-2921) with the already advanced cur_len; the loop ends when no utterance can improve
        or every candidate hit the stopping criteria (:2923-2943)
    Returns int64[B, max_new] (best finished hypothesis, prompt and EOS stripped, right-padded with eot) and lengths."""
    B = enc.shape[0]
    P = len(prompt)
    K, keep = beams, 2 * beams
    max_length = P + max_new
    xkv1 = cross_kv(sd, dims, enc)
    xkv = [(k.repeat_interleave(K, 0), v.repeat_interleave(K, 0)) for k, v in xkv1]
    self_kv = [None] * dims.n_dec
    running = torch.full((B, K, max_length), eot, dtype=torch.long)
    running[:, :, :P] = torch.tensor(list(prompt), dtype=torch.long)
    sequences = running.clone()
    run_scores = torch.zeros(B, K)
    run_scores[:, 1:] = -1e9
    beam_scores = torch.full((B, K), -1e9)
    finished = torch.zeros(B, K, dtype=torch.bool)
    unsat = torch.ones(B, 1, dtype=torch.bool)
    top_mask = torch.cat([torch.ones(K, dtype=torch.bool), torch.zeros(keep - K, dtype=torch.bool)])
    cur_len = P
    logits = decoder_forward(sd, dims, running[:, :, :P].reshape(B * K, P), 0, xkv, self_kv)[:, -1].float()
    V = logits.shape[-1]
    while True:
        lp = torch.log_softmax(logits, dim=-1)
        if cur_len == P and len(begin_suppress):
            lp[:, list(begin_suppress)] = float("-inf")
        lp = (lp.view(B, K, V) + run_scores[:, :, None]).view(B, K * V)
        top_lp, top_idx = torch.topk(lp, keep, dim=1)
        src = top_idx // V
        tok = top_idx % V
        top_seq = torch.take_along_dim(running, src[:, :, None], dim=1)
        top_seq[:, :, cur_len] = tok
        hit = (tok == eot) | (cur_len + 1 >= max_length)
        # next running beams
        run_lp = top_lp + hit.to(torch.float32) * -1.0e9
        nxt = torch.topk(run_lp, K, dim=1)[1]
        running = torch.take_along_dim(top_seq, nxt[:, :, None], dim=1)
        run_scores = torch.take_along_dim(run_lp, nxt, dim=1)
        beam_src = torch.take_along_dim(src, nxt, dim=1)
        # finished slots
        did = hit & top_mask[None, :]
        fin_lp = top_lp / ((cur_len + 1 - P) ** length_penalty)
        fin_lp = fin_lp + (~unsat).to(torch.float32) * -1.0e9
        fin_lp = fin_lp + (~did) * -1.0e9
        m_seq = torch.cat((sequences, top_seq), dim=1)
        m_sc = torch.cat((beam_scores, fin_lp), dim=1)
        m_fin = torch.cat((finished, did), dim=1)
        sel = torch.topk(m_sc, K, dim=1)[1]
        sequences = torch.take_along_dim(m_seq, sel[:, :, None], dim=1)
        beam_scores = torch.take_along_dim(m_sc, sel, dim=1)
        finished = torch.take_along_dim(m_fin, sel, dim=1)
        cur_len += 1
        best_possible = run_scores[:, :1] / ((cur_len - P) ** length_penalty)
        worst = torch.where(finished, beam_scores.min(dim=1, keepdim=True)[0], torch.tensor(-1.0e9))
        unsat = unsat & (best_possible > worst).any(dim=-1, keepdim=True)
        if not bool(unsat.any()) or bool(hit.all()):
            break
        flat = (beam_src + torch.arange(B)[:, None] * K).view(-1)
        self_kv = [(k[flat], v[flat]) for k, v in self_kv]
        logits = decoder_forward(sd, dims, running[:, :, cur_len - 1].reshape(-1, 1), cur_len - 1, xkv, self_kv)[:, -1].float()
    out = torch.full((B, max_new), eot, dtype=torch.long)
    lens = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        best = sequences[b, 0, P:].tolist()
        n = 0
        while n < max_new and best[n] != eot:
            n += 1
        out[b, :n] = torch.tensor(best[:n], dtype=torch.long)
        lens[b] = n
    return out, lens

