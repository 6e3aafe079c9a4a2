/* libwipa — C ABI of the B200-native whisper-ipa hot path (log-mel -> encoder -> KV-cached decode -> PER).
 *
 * The reference (barathanaslan/whisper-ipa) has no FFI of its own: its two entry points call Python
 * packages (mlx_whisper, editdistance).  Each entry point below names the reference call it replaces,
 * so a maintainer can bind it from the scripts with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller (PyTorch allocation);
 *     "host" pointers are ordinary host memory.  Nothing here takes or returns a torch type.
 *   - every call takes the CUDA stream to enqueue on (a cudaStream_t passed as void*), returns
 *     0 on success or a negative WIPA_E* code, and never throws.  wipa_last_error() gives detail.
 *   - the library allocates only inside an explicit wipa_ctx (weights copy, workspaces, KV caches).
 *   - one context per GPU / process; a context is not re-entrant.
 */
#ifndef WIPA_H_
#define WIPA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WIPA_OK 0
#define WIPA_EINVAL (-1)      /* bad argument / shape */
#define WIPA_ECUDA (-2)       /* CUDA runtime or driver error */
#define WIPA_ENOMEM (-3)
#define WIPA_ESTATE (-4)      /* call out of order (e.g. decode before encode, weights missing) */
#define WIPA_EUNSUPPORTED (-5)

#define WIPA_DTYPE_F32 0      /* "fp32 path": true fp32 FMA everywhere, greedy ids bit-exact vs the oracle */
/* The half-precision path: 16-bit weights / GEMM operands / KV caches, fp32 accumulation, fp32 residual stream and
 * softmax.  A build of the library carries ONE 16-bit type ("h16" below): libwipa.so is compiled for IEEE fp16 (logits
 * within 1e-3 relative L2 of the fp32 oracle), libwipa_bf16.so (the same sources with -DWIPA_H16_BF16) for bfloat16
 * (6e-3).  wipa_h16_dtype() tells which; wipa_ctx_create refuses the other one with WIPA_EUNSUPPORTED. */
#define WIPA_DTYPE_BF16 1
#define WIPA_DTYPE_F16 2

typedef struct wipa_ctx wipa_ctx;

/* Architecture of the Whisper checkpoint (what `load_model(base_model)` fixes in
 * ref:scripts/evaluate_model.py:33-37 / ref:scripts/transcribe_single.py:12). head_dim must be 64. */
typedef struct wipa_arch {
    int32_t d_model, enc_layers, dec_layers, heads, ffn, n_mels, vocab;
    int32_t dtype;            /* WIPA_DTYPE_* */
} wipa_arch;

/* One named fp32 tensor of an HF-named Whisper state_dict, resident on the device. */
typedef struct wipa_tensor_desc {
    const char* name;         /* e.g. "model.decoder.layers.3.encoder_attn.k_proj.weight" */
    const float* data;        /* device, contiguous fp32 */
    int64_t numel;
} wipa_tensor_desc;

/* ---- context ------------------------------------------------------------------------------- */
/* Replaces: mlx_whisper.load_models.load_model + model.set_dtype (ref:scripts/evaluate_model.py:33-37). */
int wipa_ctx_create(const wipa_arch* arch, int max_batch, int max_beams, wipa_ctx** out);
/* Replaces: model.update(tree_unflatten(...)) weight overlay (ref:scripts/evaluate_model.py:58-73).
 * May be called repeatedly (base weights, then the "decoder." overlay); unknown names -> WIPA_EINVAL. */
int wipa_ctx_load_weights(wipa_ctx*, const wipa_tensor_desc* tensors, int n, void* stream);
int wipa_ctx_destroy(wipa_ctx*);

/* ---- (a) log-mel front end ------------------------------------------------------------------ */
/* Replaces: mlx_whisper.audio.log_mel_spectrogram(audio, n_mels) (ref:scripts/evaluate_model.py:189,
 * ref:scripts/transcribe_single.py:45, ref:scripts/ipa_data_loader.py:82) on already pad_or_trim'ed audio.
 * audio: device f32[B, 480000];  mel: device f32[B, n_mels, 3000] (HF layout).  Context-free. */
int wipa_logmel(const float* audio, int B, int n_mels, float* mel, void* stream);

/* Audio ingest for a whole micro-batch: interleaved PCM16 at any source rate -> mono f32 at 16 kHz, zero-padded / cut to
 * out_len samples.  Replaces: mlx_whisper.audio.load_audio (decode -> s16le -> / 32768 -> mono -> 16 kHz) followed by
 * pad_or_trim (ref:scripts/evaluate_model.py:187-188, ref:scripts/transcribe_single.py:43-44, ref:scripts/ipa_data_loader.py:48,80).
 * The rate change is scipy.signal.resample_poly's polyphase FIR (the host path's resampler), see csrc/resample.cu.
 * pcm: device int16; clip b = n_ch-interleaved frames at elements [clip_off[b], clip_off[b] + clip_frames[b] * n_ch)
 * (clip_off device int64[B], clip_frames device int32[B]); taps: device f32[n_taps] = up * firwin(20 * max(up, down) + 1,
 * 1 / max(up, down), kaiser 5.0) ({1.0} with up = down = 1 for 16 kHz sources); c0 = n_pre_remove * down - n_pre_pad of
 * scipy's upfirdn bookkeeping (whisper_ipa_b200/ingest.py computes both); out: device f32[B, out_len].  Context-free. */
int wipa_resample_pcm16(const int16_t* pcm, const long long* clip_off, const int* clip_frames, int B, int n_ch, int up, int down,
                        const float* taps, int n_taps, int c0, float* out, int out_len, void* stream);

/* ---- (b) encoder ---------------------------------------------------------------------------- */
/* Replaces: model.encoder(mel) (ref:scripts/evaluate_model.py:197, ref:scripts/transcribe_single.py:54).
 * mel: device f32[B, n_mels, 3000]; enc_out: device f32[B,1500,d] or NULL.  Also projects and stores the
 * per-layer cross-attention K/V of the B utterances inside the context, ready for wipa_decode_*. */
int wipa_encode(wipa_ctx*, const float* mel, int B, float* enc_out, void* stream);
/* Same, from an externally supplied encoder output (the scripts pass `audio_features` to decode()). */
int wipa_set_audio_features(wipa_ctx*, const float* enc_out /* device f32[B,1500,d] */, int B, void* stream);

/* ---- (c) decoder ---------------------------------------------------------------------------- */
typedef struct wipa_decode_opts {
    const int32_t* prompt;        /* host int32[P]: e.g. <|sot|><|en|><|transcribe|><|notimestamps|> */
    int32_t prompt_len;
    int32_t max_new;              /* sample_len (224 - P in the reference's DecodingOptions default) */
    int32_t eot;                  /* end-of-text id; rows that emit it are padded with it afterwards */
    const int32_t* suppress;      /* host: ids masked at every step (may be NULL) */
    int32_t n_suppress;
    const int32_t* begin_suppress;/* host: ids masked at the first sampled position only */
    int32_t n_begin_suppress;
} wipa_decode_opts;

/* Replaces: mlx_whisper.decoding.decode(model, audio_features, DecodingOptions(...)) greedy path
 * (ref:scripts/evaluate_model.py:200, ref:scripts/transcribe_single.py:55, ref:scripts/train_whisper_ipa.py:356).
 * out_ids: device int32[B, max_new] (EOT-padded), out_len: device int32[B] (tokens before EOT). */
int wipa_decode_greedy(wipa_ctx*, int B, const wipa_decode_opts*, int32_t* out_ids, int32_t* out_len, void* stream);
/* Beam search with HF `_beam_search` semantics (HF:generation/utils.py:3076-3400; early_stopping=False, do_sample=False,
 * one EOS id); 2 <= beams <= min(8, max_beams of the context).  Beams of an utterance share its cross-attention K/V; the
 * self-attention cache is never reordered (a per-beam ancestry table says which slot wrote each position).  Same
 * outputs as the greedy call: the best finished hypothesis per utterance, EOS stripped, EOT-padded. */
int wipa_decode_beam(wipa_ctx*, int B, int beams, float length_penalty, const wipa_decode_opts*,
                     int32_t* out_ids, int32_t* out_len, void* stream);
/* Stepwise decoding with the logits handed back, for everything that filters or samples between steps on the caller's
 * side: HF `generate(..., logits_processor=...)`, and the long-form branch of the reference (mlx_whisper.transcribe with
 * timestamp rules, temperature fallback and no-speech detection, ref:scripts/evaluate_model.py:112-119).
 * wipa_decode_begin resets the self-KV cache, consumes T forced tokens per row (host int32[B,T]: the prompt; rows may differ)
 * and writes the logits that follow the last one (device f32[B,V]).  wipa_decode_next appends ONE token per row (device
 * int32[B], e.g. an argmax / multinomial result that never left the GPU) and writes the next logits.  Any other decode
 * call, or a new wipa_encode / wipa_set_audio_features, closes the stepwise state. */
int wipa_decode_begin(wipa_ctx*, int B, const int32_t* tokens, int T, float* logits, void* stream);
int wipa_decode_next(wipa_ctx*, int B, const int32_t* tokens_dev, float* logits, void* stream);
/* Diagnostics for the parity tests: teacher-forced logits.  tokens: host int32[B,T]; logits: device f32[B,T,V]. */
int wipa_decode_logits(wipa_ctx*, int B, const int32_t* tokens, int T, float* logits, void* stream);

/* ---- (d) PER scorer ------------------------------------------------------------------------- */
/* Replaces: editdistance.eval(ref_phones, hyp_phones) (ref:scripts/evaluate_ipa.py:100) for N pairs at once.
 * CSR packing: pair i = ref[ref_off[i]:ref_off[i+1]] vs hyp[hyp_off[i]:hyp_off[i+1]]; all device int32.
 * dist_len: device int32[N,2] = (edit distance, ref length) — the pair layout the multi-GPU gather moves.
 * max_ref_len: an upper bound of the reference lengths (sizes the per-warp shared-memory slice); a pair whose reference is
 * longer than it is flagged with distance -1 instead of being scored (checked on the device, never out of bounds).
 * The percentage (d/len)*100.0 and mean/std stay on the host in float64 (ref:scripts/evaluate_ipa.py:103,370). */
int wipa_per_batch(const int32_t* ref, const int32_t* ref_off, const int32_t* hyp, const int32_t* hyp_off,
                   int N, int max_ref_len, int32_t* dist_len, void* stream);

/* Replaces: PFERCalculator.phone_feature_error_rate (mode 0, Hamming / 24; ref:scripts/evaluate_ipa.py:163-213) and
 * PFERCalculatorCosine.phone_feature_error_rate (mode 1; ref:scripts/evaluate_ipa.py:236-287) for N pairs at once.
 * Same CSR packing of interned phone ids as wipa_per_batch; feats: device int8[n_phones, 24] (panphon's numeric
 * features, all-zero rows for unknown phones); dist: device f64[N] = D[len_ref][len_hyp] of the feature-weighted edit
 * distance, bit-identical to the reference's float64 numpy DP.  The percentage stays on the host. */
int wipa_pfer_batch(const int32_t* ref, const int32_t* ref_off, const int32_t* hyp, const int32_t* hyp_off,
                    int N, int max_ref_len, const int8_t* feats, int mode, double* dist, void* stream);

/* ---- introspection -------------------------------------------------------------------------- */
/* WIPA_DTYPE_F16 or WIPA_DTYPE_BF16: the 16-bit element type ("h16") this build of the library computes in. */
int wipa_h16_dtype(void);
const char* wipa_strerror(int code);
const char* wipa_last_error(void);
/* Number of kernels of this library launched since the last reset (bench.py's "gpu_launches"). */
int64_t wipa_launch_count(int reset);
/* Timing of one named kernel family on its own launch stream: bench.py brackets with these. */
int wipa_ctx_get_info(wipa_ctx*, int what, int64_t* out);
#define WIPA_INFO_WORKSPACE_BYTES 0
#define WIPA_INFO_CROSSKV_BYTES 1
#define WIPA_INFO_DECODE_STEPS 2
#define WIPA_INFO_XATTN_LATENT 3   /* 1: cross-attention runs over the encoder output itself (no per-layer cross-KV) */

/* Standalone kernel entry points used by tests/ and bench.py's roofline section. */
/* C[M,N] = A[M,K] * W[N,K]^T (+bias) in h16 on tcgen05, fp32 accumulate, fp32 out. All device pointers. */
int wipa_test_gemm_h16(const void* A_h16, const void* W_h16, const float* bias, float* C,
                        int M, int N, int K, int block_n, void* stream);
int wipa_test_gemm_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, void* stream);
/* Epilogue variants of the h16 GEMMs on encoder-shaped problems (M = rows_per_batch * n_batch rows, tiles never straddle
 * a batch; block_n = 0 selects the persistent kernel and its specialised epilogues).  mode: 0 bias, 1 bias + GELU,
 * 2 bias + fp32 residual, 3 bias + q|k|v head split into [3][n_batch][H][rows_per_batch][64] with N = 3*H*64
 * (HF:models/whisper/modeling_whisper.py WhisperEncoderLayer: the projections, fc1 + activation_fn, the two residual adds).
 * `out` holds h16 when out_h16 != 0, fp32 otherwise. */
int wipa_test_gemm_epilogue(const void* A, const void* W, const float* bias, const float* resid, void* out,
                            int rows_per_batch, int n_batch, int N, int K, int mode, int out_h16, int block_n, void* stream);
/* conv1d-as-GEMM addressing: logical row (batch, t) of A starts at A + batch*bstride + t*lda and spans K >= lda
 * elements (rows overlap); A/W are h16 (tcgen05 kernel) when is_h16 else fp32 (SIMT kernel); C fp32 [M, N]. */
int wipa_test_gemm_rows(const void* A, int is_h16, long long lda, int rows_per_batch, long long bstride, int n_batch,
                        const void* W, float* C, int N, int K, int block_n, void* stream);
/* Encoder self-attention alone: q,k,v device f32 [B,H,T,64] (q pre-scaled) -> out device f32 [B,T,H*64]. */
int wipa_test_enc_attention(const float* q, const float* k, const float* v, float* out, int B, int H, int T,
                            int use_h16, void* stream);
/* Same on h16 device buffers without conversions (kernel timing): tc = 1 tcgen05 kernel, 0 SIMT kernel. */
int wipa_test_enc_attention_h16(const void* q, const void* k, const void* v, void* out, int B, int H, int T, int tc, void* stream);
/* Latent cross-attention kernel alone (the decoder attends over the encoder output itself; k / v projections are folded
 * into the query / output projections, HF:models/whisper/modeling_whisper.py:241-357 WhisperAttention as cross-attention):
 * Qp h16 [S, H, 64*H] absorbed queries, E h16 [U, T, 64*H] encoder output, utt_of_seq int32 [S] -> C h16 [S, H, 64*H]
 * = softmax_t(Qp[s,h] . E[u,t]) E[u].  H = 6, 8, 12 or 16 (Whisper tiny .. medium), or 20 (large*: layout 1 or 2 only, two
 * CTAs of 10 heads per key range, csrc/attn_lat_wide.cu).  All device pointers. */
int wipa_test_cross_attn_latent(const void* Qp, const void* E, int U, const int* utt_of_seq, void* C, int S, int H, int T,
                                int layout, int beams, void* stream);
/* beams: 1, or the beam count K (2..8) when utt_of_seq[s] = s / K (the K beams of an utterance are adjacent sequences): a group of
 * K CTAs then walks each (utterance, key range) together so that E leaves HBM once per utterance, not once per beam. */
/* layout of E above: 0 = row-major, fetched as TMA boxes; 1 = row-major in, converted to the chunk-tiled layout inside the
 * call (not for timing); 2 = already chunk-tiled ([U][chunk][column tile][key][64 swizzled], made by wipa_test_lat_tile into a
 * zeroed buffer of U * wipa_test_lat_tiled_elems(H, T) h16 elements): what the context keeps and bench.py times. */
long long wipa_test_lat_tiled_elems(int H, int T);
int wipa_test_lat_tile(const void* E, int U, int T, int H, void* out, void* stream);
/* The absorbed queries of the latent cross-attention in one kernel (ctx.cu decode_step runs this per layer; the q rows stay
 * in shared memory between the two products): A h16 [S, 64*H] (LayerNorm output), Wq / Wk h16 [64*H, 64*H] row-major [out, in]
 * (HF q_proj / k_proj of WhisperAttention, modeling_whisper.py:241-357), bias f32 [64*H] or NULL, wkt_scratch h16 [H, 64*H, 64]
 * -> out h16 [S, H, 64*H], out[s, h, :] = Wk_h^T (Wq_h A[s] + bias_h).  H even.  All device pointers. */
int wipa_test_xlq_fused(const void* A, const void* Wq, const void* Wk, const float* bias, void* wkt_scratch, void* out, int S,
                        int H, void* stream);
/* One decode-step self-attention over a caller-built paged KV cache: kpool / vpool [page][H][16][64] (h16 when is_h16,
 * else f32), block_table int32 [B, bt_stride] page ids, *pos_ptr = newest position (length - 1); q f32 [B, H*64];
 * out [B, H*64] in the pool's element type.  All device pointers. */
int wipa_test_self_attn(const float* q, const void* kpool, const void* vpool, const int* block_table, int bt_stride,
                        const int* pos_ptr, void* out, int B, int H, int is_h16, void* stream);
/* The node that ends a decode step, alone: fused vocabulary projection + masked argmax over the S rows in the context's
 * LayerNorm-output buffer (weights V x d streamed once, logits never stored).  16-bit contexts. */
int wipa_test_logits_argmax(wipa_ctx*, int S, void* stream);
/* One decode-step cross-attention sweep over the context's cached encoder K/V (the dominant HBM kernel):
 * q: device f32[B, d] (pre-scaled), out: device f32[B, d]; layer selects which cached K/V. */
int wipa_test_cross_attn(wipa_ctx*, int B, int layer, const float* q, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WIPA_H_ */
