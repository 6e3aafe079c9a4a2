// wipa_ctx: weights, workspaces, cross-/self-KV caches and the encoder / decoder schedules built from the kernels.
//
// HBM layout (T = float on the fp32 path, h16 on the h16 path; residual stream always fp32):
//   weights      one arena of T (GEMM operands, [N, K] K-major, q/k/v fused to [3d, d] with q pre-scaled by the
//                exact power of two 64^-0.5) + one fp32 arena (biases, LayerNorm, position tables)
//   cross-KV     [dec layer][K|V][utterance][head][1500][64] T — one contiguous 1500x64 block per (utt, head), the
//                unit the cross-attention kernel streams with cp.async.bulk; indexed by utterance, so beams share it
//   self-KV      per layer a pool of pages [page][head][16][64] T addressed through a block table [seq][28]
//   activations  encoder micro-batch buffers sized for enc_mb clips; decoder buffers sized for max_batch*max_beams rows
#include <math.h>
#include <string.h>

#include <string>
#include <array>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "logmel.cuh"

namespace {

enum SlotKind { SLOT_F32 = 0, SLOT_T = 1, SLOT_CONV = 2 };
struct Slot {
    void* dst;
    int kind;
    float scale;
    int64_t numel;
    int N, C;            // SLOT_CONV
    std::string canon;   // canonical name (aliases share it)
    float* master;       // fp32 copy (scaled) kept for the folded-LayerNorm weights: W * gain is then rounded ONCE
};

struct EncLayer {
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    void *qkv_w, *o_w, *fc1_w, *fc2_w;
    float *qkv_b, *o_b, *fc1_b, *fc2_b;
};
struct DecLayer {
    float *qkv_m, *cq_m, *fc1_m;        // fp32 masters (folded-LayerNorm path), else null
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
    void *qkv_w, *o_w, *cq_w, *co_w, *fc1_w, *fc2_w;
    float *qkv_b, *o_b, *cq_b, *co_b, *fc1_b, *fc2_b;
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int nodes = 0;
};

}  // namespace

struct wipa_ctx {
    wipa_arch a;
    int max_batch, max_beams, max_seqs;
    bool bf;
    size_t esz;
    int enc_mb;
    int pages_per_seq;

    std::vector<void*> allocs;
    std::unordered_map<std::string, Slot> slots;
    std::unordered_set<std::string> loaded;
    size_t n_required = 0;

    // weights
    void *conv1_w, *conv2_w, *tok_emb, *xkv_w;
    float *conv1_b, *conv2_b, *enc_pos, *enc_ln_w, *enc_ln_b, *dec_pos, *dec_ln_w, *dec_ln_b, *xkv_b;
    std::vector<EncLayer> enc;
    std::vector<DecLayer> dec;

    // encoder workspaces
    void *mel_rows, *conv1_out, *eh, *eqkv, *eattn, *effn, *enc_T;
    float* ex;
    // caches
    void* xkv;
    size_t xkv_which_stride;       // elements between consecutive (layer, K|V) blocks
    int n_utts = 0;
    void *kpool, *vpool;
    size_t pool_layer_stride;      // elements per layer pool
    // decoder workspaces
    float *dx, *dq, *ca_part, *pmax, *logits;
    void *dh, *dattn, *dffn;
    float* sk_part = nullptr;      // split-K scratch of the decode fc2 GEMM
    int* sk_count = nullptr;
    int *block_table, *utt_of_seq, *ca_counters, *pidx;
    int *d_pos, *d_step, *d_cur_tok, *d_done, *d_n_done, *d_forced, *d_out_ids, *d_out_len;
    uint32_t *mask_always, *mask_begin;
    // beam search state (allocated when max_beams > 1)
    int *b_flip = nullptr, *b_run_seq = nullptr, *b_fin_seq = nullptr, *b_anc = nullptr, *b_fin_done = nullptr, *b_fin_done_next = nullptr,
        *b_unsat = nullptr, *b_cand_idx = nullptr, *b_prompt = nullptr;
    float *b_run_score = nullptr, *b_run_score_next = nullptr, *b_fin_score = nullptr, *b_fin_score_next = nullptr, *b_cand_val = nullptr;
    int* h_pinned = nullptr;       // pinned host scratch (n_done)
    cudaStream_t cap_stream = nullptr;   // graphs are captured here (the caller's stream may be the legacy default stream)
    int n_logit_tiles;
    int bn_enc, bn_dec, bn_logits, ca_split;
    int splitk = 1;                // WIPA_SPLITK=0 disables the split-K decode fc2
    int sk_bn = 32, sk_splits = 6;
    int bn_qkv_wide = 64, bn_fc1_wide = 64, bn_xlq2_wide = 128, sk_wide = 0;
    int sk_many = 0;               // WIPA_SPLITK_MANY=1: keep split-K on the long-K nodes when S > 256 (measured slower there)   // tile widths once S needs more than two M tiles (WIPA_BN_*_WIDE) // WIPA_SK_BN / WIPA_SK_SPLITS: tile width and K splits of the long-K decode GEMMs once S needs two M tiles
    int beam_L = 0;                // row length of the beam-search sequence / ancestry arrays of the current decode
    int persistent_min_tiles = 296; // WIPA_PERSISTENT_MIN_TILES: fewer 128x256 tiles than this -> plain 128x128-tile kernel
    int skip_mask = 0;             // WIPA_SKIP_MASK (timing ablation only, results become garbage): see decode_step
    int enc_attn_simt = 0;         // WIPA_ENC_ATTN_SIMT=1: SIMT flash kernel instead of the tcgen05 one (h16 path)
    // latent cross-attention (attn_lat.cu): the decoder attends over the encoder output itself, k / v projections folded
    // into the query and output projections.  Default on the h16 path for contexts of >= 96 sequences and <= 16 heads.
    int xlat = 0;
    int xl_q2 = 1;                 // WIPA_XL_Q2STEP: absorbed queries in two steps (q = x Wq^T, then q'_h = Wk_h^T q_h per head: 12 x
                                   // fewer weight bytes and 6 x fewer FLOPs than the one-step GEMM against the folded [H*d, d] matrix)
    int xl_o2 = 1;                 // WIPA_XL_O2STEP: context rows in two steps as well (ctx_h = Wv_h c_h + bv_h head-batched, then the ordinary
                                   // out-projection): no folded [d, H*d] matrix, no K = H*d split-K node
    int xl_qfused = 0;             // WIPA_XL_QFUSED: the two steps above in one kernel (gemm_q2.cu), q never leaves the SM
    std::vector<void*> xl_wkt;     // per layer [H][d][64]: Wk transposed per head (the per-head GEMM's K-major W operand)
    void* dq16 = nullptr;          // q rows [S, d] in h16 between the two steps
    int cur_beams = 1;             // beams of the decode in progress (decode_setup): the latent kernel groups an utterance's beams
    int xl_wide = 0;               // 16 / 20 heads: attn_lat_wide.cu (128 columns per warp; 20 heads as two CTAs of 10) instead of attn_lat.cu
    int xl_tiled = 1;              // WIPA_XL_TILED: the encoder output is kept chunk-tiled / pre-swizzled (bulk copies) instead of row-major (TMA boxes)
    int bn_xlq = 0;                // WIPA_BN_XLQ: tile width of the absorbed-query GEMM (0: as fc1)
    bool xlat_ready = false;       // folded weights match the loaded weights
    void *enc_lat = nullptr, *dqlat = nullptr, *dclat = nullptr;     // E [max_batch, 1500, d]; Q' and C [S, H*d]
    float* xl_part = nullptr;      // partials of sequences cut by the stream-K ranges
    size_t xl_part_floats = 0;
    std::vector<void*> xlq_w, xlo_w;                                  // per layer [H*d, d] and [d, H*d]
    std::vector<float*> xlq_b, xlo_b;
    // folded LayerNorm (common.cuh): the decode step of the 16-bit path runs without LayerNorm kernels.  Gain-folded copies of
    // the weights that consume a LayerNorm output, their column sums c and the beta-folded biases b'; the residual stream
    // rounded to h16 and its per-piece statistics.  WIPA_LN_FOLD=0 restores the LayerNorm kernels.
    int lnf = 0;
    bool lnf_ready = false;
    void *dx16 = nullptr, *emb_wf = nullptr;
    float* emb_master = nullptr;
    float *dstats = nullptr, *emb_c = nullptr, *emb_bf = nullptr;
    std::vector<void*> qkv_wf, cq_wf, fc1_wf;
    std::vector<float*> qkv_c, qkv_bf, cq_c, cq_bf, fc1_c, fc1_bf, xlq_c, xlq_beff;
    int64_t decode_steps = 0;
    int step_pos = 0;              // target positions consumed by wipa_decode_begin / wipa_decode_next (0: no stepwise decode open)
    size_t workspace_bytes = 0, xkv_bytes = 0;
    LogmelTables mel_tables;
    float* mel_clipmax;

    // captured decode steps, keyed by EVERYTHING the captured kernels bake in as parameters:
    // {kind (0 greedy, 1 beam), sequences, prompt length, max_new, eot, beams, length-penalty bits}
    std::map<std::array<long long, 7>, GraphEntry> graphs;
};

namespace {

int ctx_alloc(wipa_ctx* c, void** p, size_t bytes, bool zero) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        wipa_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return WIPA_ENOMEM;
    }
    c->allocs.push_back(*p);
    c->workspace_bytes += bytes;
    if (zero) WIPA_CUDA_CHECK(cudaMemset(*p, 0, bytes));
    return WIPA_OK;
}

struct Carver {
    char* base;
    size_t off = 0;
    size_t esz;
    void* take(size_t elems) {
        void* p = base ? base + off : nullptr;
        off += ((elems * esz + 255) & ~(size_t)255);
        return p;
    }
};

// Lays out both arenas and the name table.  Called twice: once with null bases to measure, once to assign.
void layout_weights(wipa_ctx* c, char* tbase, char* fbase, char* mbase, size_t* tbytes, size_t* fbytes, size_t* mbytes) {
    const wipa_arch& a = c->a;
    const int d = a.d_model, ffn = a.ffn, V = a.vocab;
    Carver T{tbase, 0, c->esz}, F{fbase, 0, 4}, Mst{mbase, 0, 4};
    const bool assign = tbase != nullptr;
    float* next_master = nullptr;       // set right before a reg() whose tensor keeps an fp32 master copy
    auto reg = [&](const std::string& name, void* dst, int kind, int64_t numel, float scale = 1.f, int N = 0, int C = 0,
                   const char* canon = nullptr) {
        float* m = next_master;
        next_master = nullptr;
        if (!assign) return;
        Slot s{dst, kind, scale, numel, N, C, canon ? std::string(canon) : name, m};
        c->slots[name] = s;
    };
    // fp32 masters of the weights that consume a LayerNorm output (decoder qkv, cross-attention q, fc1, the tied embedding):
    // the folded-LayerNorm path multiplies them by the gain and rounds to h16 once (ln_fold_rows)
    auto master = [&](size_t elems) -> float* { return c->lnf ? (float*)Mst.take(elems) : nullptr; };
    auto ln = [&](const std::string& p, float*& w, float*& b) {
        w = (float*)F.take(d); b = (float*)F.take(d);
        reg(p + ".weight", w, SLOT_F32, d); reg(p + ".bias", b, SLOT_F32, d);
    };
    auto lin = [&](const std::string& p, void*& w, float*& b, int N, int K) {
        w = T.take((size_t)N * K); b = (float*)F.take(N);
        reg(p + ".weight", w, SLOT_T, (int64_t)N * K); reg(p + ".bias", b, SLOT_F32, N);
    };
    // q|k|v fused: q (and its bias) carry the 64^-0.5 = 0.125 scaling (exact), k has no bias (stays zero)
    auto qkv = [&](const std::string& p, void*& w, float*& b, float* m = nullptr) {
        w = T.take((size_t)3 * d * d); b = (float*)F.take(3 * d);
        char* wb = (char*)w;
        const size_t blk = (size_t)d * d * c->esz;
        next_master = m;
        reg(p + ".q_proj.weight", wb, SLOT_T, (int64_t)d * d, 0.125f);
        next_master = m ? m + (size_t)d * d : nullptr;
        reg(p + ".k_proj.weight", wb ? wb + blk : nullptr, SLOT_T, (int64_t)d * d);
        next_master = m ? m + (size_t)2 * d * d : nullptr;
        reg(p + ".v_proj.weight", wb ? wb + 2 * blk : nullptr, SLOT_T, (int64_t)d * d);
        reg(p + ".q_proj.bias", b, SLOT_F32, d, 0.125f);
        reg(p + ".v_proj.bias", b ? b + 2 * d : nullptr, SLOT_F32, d);
    };
    if (assign) { c->enc.resize(a.enc_layers); c->dec.resize(a.dec_layers); }
    std::vector<EncLayer> enc_tmp(a.enc_layers);
    std::vector<DecLayer> dec_tmp(a.dec_layers);
    std::vector<EncLayer>& E = assign ? c->enc : enc_tmp;
    std::vector<DecLayer>& D = assign ? c->dec : dec_tmp;

    const std::string pe = "model.encoder.", pd = "model.decoder.";
    c->conv1_w = T.take((size_t)d * 3 * a.n_mels); c->conv1_b = (float*)F.take(d);
    reg(pe + "conv1.weight", c->conv1_w, SLOT_CONV, (int64_t)d * 3 * a.n_mels, 1.f, d, a.n_mels);
    reg(pe + "conv1.bias", c->conv1_b, SLOT_F32, d);
    c->conv2_w = T.take((size_t)d * 3 * d); c->conv2_b = (float*)F.take(d);
    reg(pe + "conv2.weight", c->conv2_w, SLOT_CONV, (int64_t)d * 3 * d, 1.f, d, d);
    reg(pe + "conv2.bias", c->conv2_b, SLOT_F32, d);
    c->enc_pos = (float*)F.take((size_t)WIPA_T_ENC * d);
    reg(pe + "embed_positions.weight", c->enc_pos, SLOT_F32, (int64_t)WIPA_T_ENC * d);
    for (int l = 0; l < a.enc_layers; ++l) {
        const std::string p = pe + "layers." + std::to_string(l) + ".";
        ln(p + "self_attn_layer_norm", E[l].ln1_w, E[l].ln1_b);
        qkv(p + "self_attn", E[l].qkv_w, E[l].qkv_b);
        lin(p + "self_attn.out_proj", E[l].o_w, E[l].o_b, d, d);
        ln(p + "final_layer_norm", E[l].ln2_w, E[l].ln2_b);
        lin(p + "fc1", E[l].fc1_w, E[l].fc1_b, ffn, d);
        lin(p + "fc2", E[l].fc2_w, E[l].fc2_b, d, ffn);
    }
    ln(pe + "layer_norm", c->enc_ln_w, c->enc_ln_b);

    c->tok_emb = T.take((size_t)V * d);
    c->emb_master = master((size_t)V * d);
    next_master = c->emb_master;
    reg(pd + "embed_tokens.weight", c->tok_emb, SLOT_T, (int64_t)V * d);
    next_master = c->emb_master;
    reg("proj_out.weight", c->tok_emb, SLOT_T, (int64_t)V * d, 1.f, 0, 0, "model.decoder.embed_tokens.weight");   // tied head
    c->dec_pos = (float*)F.take((size_t)WIPA_MAX_TGT * d);
    reg(pd + "embed_positions.weight", c->dec_pos, SLOT_F32, (int64_t)WIPA_MAX_TGT * d);
    c->xkv_w = T.take((size_t)a.dec_layers * 2 * d * d);
    c->xkv_b = (float*)F.take((size_t)a.dec_layers * 2 * d);
    for (int l = 0; l < a.dec_layers; ++l) {
        const std::string p = pd + "layers." + std::to_string(l) + ".";
        ln(p + "self_attn_layer_norm", D[l].ln1_w, D[l].ln1_b);
        D[l].qkv_m = master((size_t)3 * d * d);
        qkv(p + "self_attn", D[l].qkv_w, D[l].qkv_b, D[l].qkv_m);
        lin(p + "self_attn.out_proj", D[l].o_w, D[l].o_b, d, d);
        ln(p + "encoder_attn_layer_norm", D[l].ln2_w, D[l].ln2_b);
        // cross-attention: q scaled like self-attention; k / v of all layers fused into one [L*2*d, d] operand
        D[l].cq_w = T.take((size_t)d * d); D[l].cq_b = (float*)F.take(d);
        D[l].cq_m = (c->lnf && (!c->xlat || c->xl_q2)) ? master((size_t)d * d) : nullptr;
        next_master = D[l].cq_m;
        reg(p + "encoder_attn.q_proj.weight", D[l].cq_w, SLOT_T, (int64_t)d * d, 0.125f);
        reg(p + "encoder_attn.q_proj.bias", D[l].cq_b, SLOT_F32, d, 0.125f);
        char* xw = (char*)c->xkv_w;
        const size_t blk = (size_t)d * d * c->esz;
        reg(p + "encoder_attn.k_proj.weight", xw ? xw + (size_t)(2 * l) * blk : nullptr, SLOT_T, (int64_t)d * d);
        reg(p + "encoder_attn.v_proj.weight", xw ? xw + (size_t)(2 * l + 1) * blk : nullptr, SLOT_T, (int64_t)d * d);
        reg(p + "encoder_attn.v_proj.bias", c->xkv_b ? c->xkv_b + (size_t)(2 * l + 1) * d : nullptr, SLOT_F32, d);
        lin(p + "encoder_attn.out_proj", D[l].co_w, D[l].co_b, d, d);
        ln(p + "final_layer_norm", D[l].ln3_w, D[l].ln3_b);
        D[l].fc1_m = master((size_t)ffn * d);
        next_master = D[l].fc1_m;
        lin(p + "fc1", D[l].fc1_w, D[l].fc1_b, ffn, d);
        lin(p + "fc2", D[l].fc2_w, D[l].fc2_b, d, ffn);
    }
    ln(pd + "layer_norm", c->dec_ln_w, c->dec_ln_b);
    *tbytes = T.off;
    *fbytes = F.off;
    *mbytes = Mst.off;
}

AOperand plainA(const void* p, int M, int K) {
    AOperand a;
    a.ptr = p; a.lda = K; a.a_rpb = M; a.a_bstride = (long long)M * K; a.n_batch = 1;
    return a;
}

EpiParams epi(int mode, int M, int N) {
    EpiParams e;
    memset(&e, 0, sizeof(e));
    e.mode = mode;
    e.N = N;
    e.vec_ok = (N % 8 == 0) ? 1 : 0;
    e.o_rpb = M > 0 ? M : 1;
    e.ldo = N;
    e.vt_which = -1;
    e.T = 1;
    return e;
}

int gemm(wipa_ctx* c, const AOperand& a, const void* W, int M, int N, int K, const EpiParams& ep, int bn, cudaStream_t st) {
    if (c->bf) {
        // bn == 0: the persistent 128 x 256 kernel when there are at least two waves of its tiles, else 128 x 128 tiles
        // (the vocabulary argmax always takes the persistent kernel: its 128 x 256 tiles stream the embedding matrix once)
        if (bn == 0 && ep.mode != EPI_ARGMAX && (long long)cdiv(M, 128) * cdiv(N, 256) < c->persistent_min_tiles) bn = 128;
        return launch_gemm_h16(a, (const h16*)W, M, N, K, ep, bn, st);
    }
    return launch_gemm_f32(a, (const float*)W, M, N, K, ep, st);
}

template <typename T>
int ln_t(const float* x, const float* w, const float* b, void* out, int M, int d, cudaStream_t st) {
    return launch_layernorm<T>(x, w, b, (T*)out, M, d, st);
}
int ln(wipa_ctx* c, const float* x, const float* w, const float* b, void* out, int M, cudaStream_t st) {
    return c->bf ? ln_t<h16>(x, w, b, out, M, c->a.d_model, st) : ln_t<float>(x, w, b, out, M, c->a.d_model, st);
}

__global__ void to_f32_kernel(const void* __restrict__ src, int is_h16, float* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        dst[i] = is_h16 ? h16_to_f32(reinterpret_cast<const h16*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
}
int launch_to_f32(const void* src, int is_h16, float* dst, long long n, cudaStream_t st) {
    if (n == 0) return WIPA_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    to_f32_kernel<<<(int)blocks, 256, 0, st>>>(src, is_h16, dst, n);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

__global__ void decode_init_kernel(DecodeState ds, int* block_table, int* utt_of_seq, int Bs, int beams, int pages_per_seq,
                                   int* ca_counters, int n_counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        *ds.pos = 0;
        *ds.step = -(ds.n_forced - 1);
        *ds.n_done = 0;
        *ds.ticket = 0;
    }
    if (i < Bs) {
        ds.cur_tok[i] = ds.forced[(size_t)i * ds.n_forced];
        ds.done[i] = 0;
        ds.out_len[i] = ds.max_new;
        utt_of_seq[i] = i / beams;
    }
    if (i < Bs * pages_per_seq) block_table[i] = i;          // sequence b owns pages [b*pps, (b+1)*pps)
    if (i < Bs * ds.max_new) ds.out_ids[i] = ds.eot;
    if (i < n_counters) ca_counters[i] = 0;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int require_weights(wipa_ctx* c) {
    WIPA_CHECK(c->loaded.size() >= c->n_required, WIPA_ESTATE, "weights missing: %zu of %zu tensors loaded", c->loaded.size(),
               c->n_required);
    return WIPA_OK;
}

// ------------------------------------------------------------------------------------------------
// latent cross-attention: fold Wk into the query projection and Wv into the output projection (once per weight load)
//   Wq'[(h, n), k] = sum_j Wk[64h + j, n] Wq[64h + j, k]      bq'[(h, n)] = sum_j Wk[64h + j, n] bq[64h + j]
//   Wo'[m, (h, n)] = sum_j Wo[m, 64h + j] Wv[64h + j, n]      bo'[m]      = bo[m] + sum_i Wo[m, i] bv[i]
// (Wq / bq already carry the 1/8 scaling; k_proj has no bias.)  fp32 sums of the stored h16 weights, rounded once.
// ------------------------------------------------------------------------------------------------
// `gain` (nullable): the LayerNorm gain of the folded-LayerNorm decode path, multiplied onto input column k before the single
// rounding; `bq` is then the beta-folded bias bq + Wq beta (ln_fold_rows), which carries through the fold unchanged
__global__ void xlat_fold_q_kernel(const h16* __restrict__ Wq, const float* __restrict__ bq, const h16* __restrict__ Wk,
                                   h16* __restrict__ Wq2, float* __restrict__ bq2, int d, const float* __restrict__ gain) {
    const int h = blockIdx.z, n = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float wk[64];
    if (threadIdx.x < 64) wk[threadIdx.x] = h16_to_f32(Wk[(size_t)(h * 64 + threadIdx.x) * d + n]);
    __syncthreads();
    if (k < d) {
        float acc = 0.f;
        for (int j = 0; j < 64; ++j) acc = fmaf(wk[j], h16_to_f32(Wq[(size_t)(h * 64 + j) * d + k]), acc);
        Wq2[((size_t)h * d + n) * d + k] = f32_to_h16(gain != nullptr ? acc * gain[k] : acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float acc = 0.f;
        for (int j = 0; j < 64; ++j) acc = fmaf(wk[j], bq[h * 64 + j], acc);
        bq2[h * d + n] = acc;
    }
}

__global__ void xlat_fold_o_kernel(const h16* __restrict__ Wo, const float* __restrict__ bo, const h16* __restrict__ Wv,
                                   const float* __restrict__ bv, h16* __restrict__ Wo2, float* __restrict__ bo2, int d, int H) {
    const int h = blockIdx.z, m = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float wo[64];
    if (threadIdx.x < 64) wo[threadIdx.x] = h16_to_f32(Wo[(size_t)m * d + h * 64 + threadIdx.x]);
    __syncthreads();
    if (n < d) {
        float acc = 0.f;
        for (int j = 0; j < 64; ++j) acc = fmaf(wo[j], h16_to_f32(Wv[(size_t)(h * 64 + j) * d + n]), acc);
        Wo2[(size_t)m * ((size_t)H * d) + (size_t)h * d + n] = f32_to_h16(acc);
    }
    if (h == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        float acc = bo[m];
        for (int i = 0; i < d; ++i) acc = fmaf(h16_to_f32(Wo[(size_t)m * d + i]), bv[i], acc);
        bo2[m] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// folded LayerNorm: W'[n, k] = W[n, k] * gain[k] (one rounding), c[n] = sum_k W'[n, k] (of the ROUNDED values: what the
// tensor cores will multiply), b'[n] = b[n] + sum_k beta[k] W[n, k].  One warp per output row; any output may be null.
// ------------------------------------------------------------------------------------------------
// W is the fp32 master when `Wm` is given (gain-folded weights are then rounded once), else the stored h16 weights.
__global__ void __launch_bounds__(256)
ln_fold_rows_kernel(const h16* __restrict__ W, const float* __restrict__ Wm, const float* __restrict__ gain, const float* __restrict__ beta,
                    const float* __restrict__ bias, h16* __restrict__ Wf, float* __restrict__ csum, float* __restrict__ bf, int N, int K) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float cs = 0.f, bs = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = Wm != nullptr ? Wm[(size_t)n * K + k] : h16_to_f32(W[(size_t)n * K + k]);
        const h16 wf = f32_to_h16(gain != nullptr ? w * gain[k] : w);
        if (Wf != nullptr) Wf[(size_t)n * K + k] = wf;
        cs += h16_to_f32(wf);
        if (beta != nullptr) bs = fmaf(beta[k], w, bs);
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) {
        if (csum != nullptr) csum[n] = cs;
        if (bf != nullptr) bf[n] = (bias != nullptr ? bias[n] : 0.f) + bs;
    }
}

int ln_fold_rows(const void* W, const float* gain, const float* beta, const float* bias, void* Wf, float* csum, float* bf, int N, int K,
                 cudaStream_t st, const float* Wm = nullptr) {
    ln_fold_rows_kernel<<<cdiv(N, 8), 256, 0, st>>>((const h16*)W, Wm, gain, beta, bias, (h16*)Wf, csum, bf, N, K);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// WkT[h][n][j] = Wk[64 h + j][n]: the K-major W operand of the per-head step q'_h = Wk_h^T q_h (K = 64)
__global__ void xlat_wkt_kernel(const h16* __restrict__ Wk, h16* __restrict__ WkT, int d) {
    const int h = blockIdx.y, n = blockIdx.x * 4 + (threadIdx.x >> 6), j = threadIdx.x & 63;
    if (n < d) WkT[((size_t)h * d + n) * 64 + j] = Wk[(size_t)(h * 64 + j) * d + n];
}

int xlat_prepare(wipa_ctx* c, cudaStream_t st) {
    if (!c->xlat || c->xlat_ready) return WIPA_OK;
    const int d = c->a.d_model, H = c->a.heads;
    const dim3 grid(cdiv(d, 256), d, H);
    for (int l = 0; l < c->a.dec_layers; ++l) {
        const DecLayer& L = c->dec[l];
        const h16* Wk = (const h16*)c->xkv_w + (size_t)(2 * l) * d * d;
        const h16* Wv = (const h16*)c->xkv_w + (size_t)(2 * l + 1) * d * d;
        const float* bv = c->xkv_b + (size_t)(2 * l + 1) * d;
        if (c->xl_q2) {
            xlat_wkt_kernel<<<dim3(cdiv(d, 4), H), 256, 0, st>>>(Wk, (h16*)c->xl_wkt[l], d);
            WIPA_LAUNCHED();
            if (!c->xl_o2) {
                xlat_fold_o_kernel<<<grid, 256, 0, st>>>((const h16*)L.co_w, L.co_b, Wv, bv, (h16*)c->xlo_w[l], c->xlo_b[l], d, H);
                WIPA_LAUNCHED();
            }
            continue;
        }
        const float* bq = L.cq_b;
        if (c->lnf) {       // beta of the cross-attention LayerNorm folds into the query bias first: bq + Wq beta
            WIPA_TRY(ln_fold_rows(L.cq_w, nullptr, L.ln2_b, L.cq_b, nullptr, nullptr, c->xlq_beff[l], d, d, st));
            bq = c->xlq_beff[l];
        }
        xlat_fold_q_kernel<<<grid, 256, 0, st>>>((const h16*)L.cq_w, bq, Wk, (h16*)c->xlq_w[l], c->xlq_b[l], d, c->lnf ? L.ln2_w : nullptr);
        WIPA_LAUNCHED();
        if (c->lnf) {       // column sums of the finished (gain-folded, rounded) weights; beta is all zero here
            WIPA_TRY(ln_fold_rows(c->xlq_w[l], nullptr, nullptr, nullptr, nullptr, c->xlq_c[l], nullptr, H * d, d, st));
        }
        if (!c->xl_o2) {
            xlat_fold_o_kernel<<<grid, 256, 0, st>>>((const h16*)L.co_w, L.co_b, Wv, bv, (h16*)c->xlo_w[l], c->xlo_b[l], d, H);
            WIPA_LAUNCHED();
        }
    }
    c->xlat_ready = true;
    return WIPA_OK;
}

// gain-folded copies of the decoder weights that consume a LayerNorm output (once per weight load)
int lnf_prepare(wipa_ctx* c, cudaStream_t st) {
    if (!c->lnf || c->lnf_ready) return WIPA_OK;
    const int d = c->a.d_model, ffn = c->a.ffn, V = c->a.vocab;
    for (int l = 0; l < c->a.dec_layers; ++l) {
        const DecLayer& L = c->dec[l];
        WIPA_TRY(ln_fold_rows(L.qkv_w, L.ln1_w, L.ln1_b, L.qkv_b, c->qkv_wf[l], c->qkv_c[l], c->qkv_bf[l], 3 * d, d, st, L.qkv_m));
        if (!c->xlat || c->xl_q2) WIPA_TRY(ln_fold_rows(L.cq_w, L.ln2_w, L.ln2_b, L.cq_b, c->cq_wf[l], c->cq_c[l], c->cq_bf[l], d, d, st, L.cq_m));
        WIPA_TRY(ln_fold_rows(L.fc1_w, L.ln3_w, L.ln3_b, L.fc1_b, c->fc1_wf[l], c->fc1_c[l], c->fc1_bf[l], ffn, d, st, L.fc1_m));
    }
    WIPA_TRY(ln_fold_rows(c->tok_emb, c->dec_ln_w, c->dec_ln_b, nullptr, c->emb_wf, c->emb_c, c->emb_bf, V, d, st, c->emb_master));
    c->lnf_ready = true;
    return WIPA_OK;
}

// ------------------------------------------------------------------------------------------------
// encoder over clips [u0, u0 + nb) of the current batch; fills enc_T rows and the cross-KV of those utterances
// ------------------------------------------------------------------------------------------------
int cross_kv_project(wipa_ctx* c, int u0, int nb, cudaStream_t st) {
    const wipa_arch& a = c->a;
    const int d = a.d_model, H = a.heads, M = nb * WIPA_T_ENC, N = a.dec_layers * 2 * d;
    EpiParams ep = epi(EPI_HEADS, M, N);
    ep.bias = c->xkv_b;
    ep.out = (char*)c->xkv + (size_t)u0 * H * WIPA_T_ENC * 64 * c->esz;
    ep.out_h16 = c->bf;
    ep.T = WIPA_T_ENC; ep.H = H; ep.d = d;
    ep.which_stride = (long long)c->xkv_which_stride;
    return gemm(c, plainA(c->enc_T, M, d), c->xkv_w, M, N, d, ep, c->bn_enc, st);
}

int encode_chunk(wipa_ctx* c, const float* mel, int u0, int nb, float* enc_out, cudaStream_t st) {
    const wipa_arch& a = c->a;
    const int d = a.d_model, H = a.heads, ffn = a.ffn, C = a.n_mels;
    const int T3 = WIPA_N_FRAMES, T = WIPA_T_ENC, M = nb * T;
    if (c->bf) WIPA_TRY(launch_mel_to_rows<h16>(mel, (h16*)c->mel_rows, nb, C, st));
    else WIPA_TRY(launch_mel_to_rows<float>(mel, (float*)c->mel_rows, nb, C, st));
    {   // conv1 (k3, p1) + GELU -> rows 1..3000 of the zero-padded [nb, 3002, d] buffer
        AOperand A; A.ptr = c->mel_rows; A.lda = C; A.a_rpb = T3; A.a_bstride = (long long)(T3 + 2) * C; A.n_batch = nb;
        EpiParams ep = epi(EPI_GELU, nb * T3, d);
        ep.gelu_fast = c->bf;
        ep.bias = c->conv1_b;
        ep.out = (char*)c->conv1_out + (size_t)d * c->esz; ep.out_h16 = c->bf;
        ep.ldo = d; ep.o_rpb = T3; ep.o_bstride = (long long)(T3 + 2) * d;
        WIPA_TRY(gemm(c, A, c->conv1_w, nb * T3, d, 3 * C, ep, c->bn_enc, st));
    }
    {   // conv2 (k3, s2, p1) + GELU + sinusoidal positions -> residual stream x f32 [nb*1500, d]
        AOperand A; A.ptr = c->conv1_out; A.lda = 2 * d; A.a_rpb = T; A.a_bstride = (long long)(T3 + 2) * d; A.n_batch = nb;
        EpiParams ep = epi(EPI_GELU_POS, M, d);
        ep.bias = c->conv2_b; ep.out = c->ex; ep.pos = c->enc_pos; ep.T = T;
        ep.ldo = d; ep.o_rpb = T; ep.o_bstride = (long long)T * d;
        WIPA_TRY(gemm(c, A, c->conv2_w, M, d, 3 * d, ep, c->bn_enc, st));
    }
    const size_t hq = (size_t)nb * H * T * 64;             // elements of one of q / k / v in [B,H,T,64]
    for (int l = 0; l < a.enc_layers; ++l) {
        const EncLayer& L = c->enc[l];
        WIPA_TRY(ln(c, c->ex, L.ln1_w, L.ln1_b, c->eh, M, st));
        {
            EpiParams ep = epi(EPI_HEADS, M, 3 * d);
            ep.bias = L.qkv_b; ep.out = c->eqkv; ep.out_h16 = c->bf;
            ep.T = T; ep.H = H; ep.d = d; ep.which_stride = (long long)hq;
            WIPA_TRY(gemm(c, plainA(c->eh, M, d), L.qkv_w, M, 3 * d, d, ep, c->bn_enc, st));
        }
        if (c->bf) {
            const h16* q = (const h16*)c->eqkv;
            if (c->enc_attn_simt) WIPA_TRY(launch_enc_attention<h16>(q, q + hq, q + 2 * hq, (h16*)c->eattn, nb, H, T, st));
            else WIPA_TRY(launch_enc_attention_tc(q, q + hq, q + 2 * hq, (h16*)c->eattn, nb, H, T, st));
        } else {
            const float* q = (const float*)c->eqkv;
            WIPA_TRY(launch_enc_attention<float>(q, q + hq, q + 2 * hq, (float*)c->eattn, nb, H, T, st));
        }
        {
            EpiParams ep = epi(EPI_RESADD, M, d);
            ep.bias = L.o_b; ep.out = c->ex; ep.resid = c->ex;
            WIPA_TRY(gemm(c, plainA(c->eattn, M, d), L.o_w, M, d, d, ep, c->bn_enc, st));
        }
        WIPA_TRY(ln(c, c->ex, L.ln2_w, L.ln2_b, c->eh, M, st));
        {
            EpiParams ep = epi(EPI_GELU, M, ffn);
            ep.bias = L.fc1_b; ep.out = c->effn; ep.out_h16 = c->bf; ep.gelu_fast = c->bf;
            WIPA_TRY(gemm(c, plainA(c->eh, M, d), L.fc1_w, M, ffn, d, ep, c->bn_enc, st));
        }
        {
            EpiParams ep = epi(EPI_RESADD, M, d);
            ep.bias = L.fc2_b; ep.out = c->ex; ep.resid = c->ex;
            WIPA_TRY(gemm(c, plainA(c->effn, M, ffn), L.fc2_w, M, d, ffn, ep, c->bn_enc, st));
        }
    }
    // latent cross-attention keeps the encoder output itself (h16) for the whole batch instead of per-layer K / V
    void* enc_dst = c->xlat ? (void*)((h16*)c->enc_lat + (size_t)u0 * T * d) : c->enc_T;
    if (c->xlat && c->xl_tiled)
        WIPA_TRY(launch_layernorm_lat(c->ex, c->enc_ln_w, c->enc_ln_b, (h16*)c->enc_lat, M, d, T, u0, cross_attention_latent_keys(H), st));
    else
    WIPA_TRY(ln(c, c->ex, c->enc_ln_w, c->enc_ln_b, enc_dst, M, st));
    if (enc_out != nullptr) WIPA_TRY(launch_layernorm<float>(c->ex, c->enc_ln_w, c->enc_ln_b, enc_out, M, d, st));
    if (c->xlat) return WIPA_OK;
    return cross_kv_project(c, u0, nb, st);
}

// ------------------------------------------------------------------------------------------------
// one decoder step for S sequences: consumes cur_tok at *pos, appends self-KV, leaves the next token in cur_tok
//   logits_mode 0: skip the vocabulary projection (teacher-forced prompt positions)
//               1: fused / fp32 argmax with the suppress masks
//               2: store fp32 logits rows at logits_out (row stride ldo)
// ------------------------------------------------------------------------------------------------
// fused vocabulary projection + running argmax over the rows in c->dh: the [S, V] logits never reach HBM;
// leaves (max, argmax) per (row, n-tile) in pmax / pidx for the finalize kernel
int logits_argmax(wipa_ctx* c, int S, int* n_tiles, cudaStream_t st) {
    EpiParams ep = epi(EPI_ARGMAX, S, c->a.vocab);
    ep.pmax = c->pmax; ep.pidx = c->pidx;
    ep.mask_always = c->mask_always; ep.mask_begin = c->mask_begin; ep.step_ptr = c->d_step;
    // bn_logits 0: persistent 128 x 256 tiles, one (max, argmax) pair per 128-column half; 128: one CTA per 128-column tile
    *n_tiles = c->bn_logits == 0 ? 2 * cdiv(c->a.vocab, 256) : cdiv(c->a.vocab, c->bn_logits);
    if (c->lnf) {       // folded final LayerNorm: raw residual (h16) x gain-folded embedding matrix, finished in the epilogue
        ep.ln_stats = c->dstats; ep.ln_c = c->emb_c; ep.ln_nt = c->a.d_model / WIPA_LN_PIECE; ep.bias = c->emb_bf;
        return gemm(c, plainA(c->dx16, S, c->a.d_model), c->emb_wf, S, c->a.vocab, c->a.d_model, ep, c->bn_logits, st);
    }
    return gemm(c, plainA(c->dh, S, c->a.d_model), c->tok_emb, S, c->a.vocab, c->a.d_model, ep, c->bn_logits, st);
}

DecodeState make_state(wipa_ctx* c, int n_forced, int max_new, int eot) {
    DecodeState ds;
    ds.pos = c->d_pos; ds.step = c->d_step; ds.cur_tok = c->d_cur_tok; ds.done = c->d_done; ds.n_done = c->d_n_done;
    ds.ticket = c->d_pos + 3;
    ds.forced = c->d_forced; ds.n_forced = n_forced; ds.out_ids = c->d_out_ids; ds.out_len = c->d_out_len;
    ds.max_new = max_new; ds.eot = eot;
    return ds;
}

int decode_step(wipa_ctx* c, int S, const DecodeState& ds, int logits_mode, float* logits_out, long long ldo, cudaStream_t st,
                bool beam = false, bool finalize = true) {
    const wipa_arch& a = c->a;
    const int d = a.d_model, H = a.heads, ffn = a.ffn, V = a.vocab;
    const int skip = c->skip_mask;       // ablation bits: 1 LN, 2 self-attn, 4 cross-attn, 8 qkv, 16 d x d GEMMs, 32 fc1, 64 fc2, 128 logits
    // folded LayerNorm (common.cuh): no LayerNorm kernels; every consumer of a LayerNorm output reads the raw residual in h16
    // (dx16) with gain-folded weights and finishes the normalisation in its epilogue from the statistics (dstats) that the
    // producer of the residual left behind
    const bool lnf = c->lnf != 0;
    const int ln_nt = d / WIPA_LN_PIECE;
    // long-K GEMMs (folded cross-attention out-projection, fc2): with two or more M tiles the 32-column tiles re-read the
    // activations 24 times per K split; 64-column tiles and six K splits halve that and still fill one wave
    const bool many = c->bf && S > 256;               // more than two M tiles: 32-column tiles would need two waves of CTAs
    const bool wide_sk = c->bf && c->splitk && d % 64 == 0 && ((S > 128 && c->sk_bn == 64) || (many && c->sk_wide));
    const int sk_splits = (many && c->sk_wide && c->sk_bn != 64) ? 3 : c->sk_splits;
    const int sk_bn = wide_sk ? 64 : c->bn_dec;
    const int bn_qkv = many ? c->bn_qkv_wide : c->bn_dec;
    const int bn_fc1 = many ? c->bn_fc1_wide : ((S > 128 && c->bn_dec == 32) ? 64 : c->bn_dec);
    const int bn_xlq2 = many ? c->bn_xlq2_wide : 128;
    auto consume_ln = [&](EpiParams& ep, const float* csum, const float* bias_f) {
        ep.ln_stats = c->dstats; ep.ln_c = csum; ep.ln_nt = ln_nt; ep.bias = bias_f;
    };
    auto produce_ln = [&](EpiParams& ep) {
        if (lnf) { ep.x16_out = c->dx16; ep.ln_stats_out = c->dstats; }
    };
    if (lnf) WIPA_TRY(launch_embed_lnf((const h16*)c->tok_emb, c->dec_pos, c->d_cur_tok, c->d_pos, c->dx, (h16*)c->dx16, c->dstats, S, d, st));
    else if (c->bf) WIPA_TRY(launch_embed<h16>((const h16*)c->tok_emb, c->dec_pos, c->d_cur_tok, c->d_pos, c->dx, S, d, st));
    else WIPA_TRY(launch_embed<float>((const float*)c->tok_emb, c->dec_pos, c->d_cur_tok, c->d_pos, c->dx, S, d, st));
    for (int l = 0; l < a.dec_layers; ++l) {
        const DecLayer& L = c->dec[l];
        char* kp = (char*)c->kpool + (size_t)l * c->pool_layer_stride * c->esz;
        char* vp = (char*)c->vpool + (size_t)l * c->pool_layer_stride * c->esz;
        if (!lnf && !(skip & 1)) WIPA_TRY(ln(c, c->dx, L.ln1_w, L.ln1_b, c->dh, S, st));
        {
            EpiParams ep = epi(EPI_QKV_DEC, S, 3 * d);
            ep.bias = L.qkv_b; ep.out = c->dq; ep.out1 = kp; ep.out2 = vp; ep.out_h16 = c->bf;
            ep.H = H; ep.d = d; ep.pos_ptr = c->d_pos; ep.block_table = c->block_table; ep.bt_stride = c->pages_per_seq;
            if (lnf) consume_ln(ep, c->qkv_c[l], c->qkv_bf[l]);
            if (!(skip & 8)) WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), lnf ? c->qkv_wf[l] : L.qkv_w, S, 3 * d, d, ep, bn_qkv, st));
        }
        const int* anc = beam ? c->b_anc : nullptr;            // beam search reads every position from the slot that wrote it
        if (skip & 2) {}
        else if (c->bf) WIPA_TRY(launch_self_attention<h16>(c->dq, (const h16*)kp, (const h16*)vp, c->block_table, c->pages_per_seq,
                                                        c->d_pos, (h16*)c->dattn, S, H, st, anc, c->b_flip, c->beam_L));
        else WIPA_TRY(launch_self_attention<float>(c->dq, (const float*)kp, (const float*)vp, c->block_table, c->pages_per_seq,
                                                   c->d_pos, (float*)c->dattn, S, H, st, anc, c->b_flip, c->beam_L));
        {
            EpiParams ep = epi(EPI_RESADD, S, d);
            ep.bias = L.o_b; ep.out = c->dx; ep.resid = c->dx;
            produce_ln(ep);
            if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(c->dattn, S, d), L.o_w, S, d, d, ep, c->bn_dec, st));
        }
        if (!lnf && !(skip & 1)) WIPA_TRY(ln(c, c->dx, L.ln2_w, L.ln2_b, c->dh, S, st));
        if (c->xlat) {
            const int Hd = H * d;
            if (c->xl_q2 && c->xl_qfused) {
                // both steps in one node: the q tile stays in shared memory (gemm_q2.cu)
                if (!(skip & 16)) WIPA_TRY(launch_xlq_fused((const h16*)(lnf ? c->dx16 : c->dh), (const h16*)(lnf ? c->cq_wf[l] : L.cq_w),
                                                            (const h16*)c->xl_wkt[l], lnf ? c->cq_bf[l] : L.cq_b, lnf ? c->dstats : nullptr,
                                                            lnf ? c->cq_c[l] : nullptr, ln_nt, (h16*)c->dqlat, S, H, st));
            } else if (c->xl_q2) {
                {   // step 1: q = LN(x) Wq^T + bq (scaled by 2^-3 in the weights) -> h16 [S, d]
                    EpiParams ep = epi(EPI_STORE, S, d);
                    ep.bias = L.cq_b; ep.out = c->dq16; ep.out_h16 = 1;
                    if (lnf) consume_ln(ep, c->cq_c[l], c->cq_bf[l]);
                    if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), lnf ? c->cq_wf[l] : L.cq_w, S, d, d, ep, c->bn_dec, st));
                }
                {   // step 2: q'_h = Wk_h^T q_h for every head: H GEMMs [S, 64] x [64, d] in one launch (batch = head; A is the
                    // 64-column slice of q, W the per-head transposed k-projection) -> h16 [S, H, d]
                    AOperand A; A.ptr = c->dq16; A.lda = d; A.a_rpb = S; A.a_bstride = WIPA_HEAD_DIM; A.n_batch = H;
                    EpiParams ep = epi(EPI_STORE, S * H, d);
                    ep.out = c->dqlat; ep.out_h16 = 1;
                    ep.o_rpb = S; ep.o_bstride = d; ep.ldo = Hd; ep.w_brows = d;
                    if (!(skip & 16)) WIPA_TRY(gemm(c, A, c->xl_wkt[l], S * H, d, WIPA_HEAD_DIM, ep, bn_xlq2, st));
                }
            } else
            {   // q' = LN(x) Wq'^T + bq'  -> h16 [S, H, d]
                EpiParams ep = epi(EPI_STORE, S, Hd);
                ep.bias = c->xlq_b[l]; ep.out = c->dqlat; ep.out_h16 = 1;
                if (lnf) consume_ln(ep, c->xlq_c[l], c->xlq_b[l]);          // xlq_w / xlq_b were folded with the gain / beta already
                if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), c->xlq_w[l], S, Hd, d, ep, c->bn_xlq ? c->bn_xlq : ((S > 128 && c->bn_dec == 32) ? 64 : c->bn_dec), st));
            }
            if (c->xl_wide) {
                if (!(skip & 4)) WIPA_TRY(launch_cross_attention_latent_wide((const h16*)c->dqlat, (const h16*)c->enc_lat, c->utt_of_seq, (h16*)c->dclat, S, H,
                                                                            WIPA_T_ENC, c->xl_part, c->xl_part_floats, c->ca_counters, st,
                                                                            beam ? c->cur_beams : 1));
            } else
            if (!(skip & 4)) WIPA_TRY(launch_cross_attention_latent((const h16*)c->dqlat, (const h16*)c->enc_lat, c->xl_tiled, c->max_batch, c->utt_of_seq,
                                                                   (h16*)c->dclat, S, H, WIPA_T_ENC, c->xl_part, c->xl_part_floats,
                                                                   c->ca_counters, st, beam ? c->cur_beams : 1));
            if (c->xl_o2) {
                {   // ctx_h = Wv_h c_h + bv_h for every head: H GEMMs [S, d] x [d, 64] in one launch (batch = head; A is head h's
                    // slice of the normalised sums, W the 64 rows of the v-projection that belong to head h) -> h16 [S, d]
                    AOperand A; A.ptr = c->dclat; A.lda = Hd; A.a_rpb = S; A.a_bstride = d; A.n_batch = H;
                    EpiParams ep = epi(EPI_STORE, S * H, WIPA_HEAD_DIM);
                    ep.bias = c->xkv_b + (size_t)(2 * l + 1) * d;
                    ep.out = c->dattn; ep.out_h16 = 1;
                    ep.o_rpb = S; ep.o_bstride = WIPA_HEAD_DIM; ep.ldo = d; ep.w_brows = WIPA_HEAD_DIM;
                    const char* Wv = (const char*)c->xkv_w + (size_t)(2 * l + 1) * d * d * c->esz;
                    if (!(skip & 16)) WIPA_TRY(gemm(c, A, Wv, S * H, WIPA_HEAD_DIM, d, ep, 32, st));
                }
                {   // x += ctx Wo^T + bo: the ordinary out-projection
                    EpiParams ep = epi(EPI_RESADD, S, d);
                    ep.bias = L.co_b; ep.out = c->dx; ep.resid = c->dx;
                    produce_ln(ep);
                    if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(c->dattn, S, d), L.co_w, S, d, d, ep, c->bn_dec, st));
                }
            } else
            {   // x += C Wo'^T + bo'   (K = H * d is long: split-K)
                EpiParams ep = epi(EPI_RESADD, S, d);
                ep.bias = c->xlo_b[l]; ep.out = c->dx; ep.resid = c->dx;
                if (c->splitk && (!many || c->sk_many)) { ep.sk_part = c->sk_part; ep.sk_count = c->sk_count; ep.sk_splits = wide_sk ? sk_splits : 0; }
                produce_ln(ep);
                if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(c->dclat, S, Hd), c->xlo_w[l], S, d, Hd, ep, sk_bn, st));
            }
        } else {
        {
            EpiParams ep = epi(EPI_STORE, S, d);
            ep.bias = L.cq_b; ep.out = c->dq; ep.out_h16 = 0;
            if (lnf) consume_ln(ep, c->cq_c[l], c->cq_bf[l]);
            if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), lnf ? c->cq_wf[l] : L.cq_w, S, d, d, ep, c->bn_dec, st));
        }
        {
            const char* xk = (const char*)c->xkv + (size_t)(2 * l) * c->xkv_which_stride * c->esz;
            const char* xv = (const char*)c->xkv + (size_t)(2 * l + 1) * c->xkv_which_stride * c->esz;
            if (skip & 4) {}
            else if (c->bf) WIPA_TRY(launch_cross_attention<h16>(c->dq, (const h16*)xk, (const h16*)xv, c->utt_of_seq, (h16*)c->dattn,
                                                             c->ca_part, c->ca_counters, S, H, c->ca_split, 1, st));
            else WIPA_TRY(launch_cross_attention<float>(c->dq, (const float*)xk, (const float*)xv, c->utt_of_seq, (float*)c->dattn,
                                                        c->ca_part, c->ca_counters, S, H, c->ca_split, 1, st));
        }
        {
            EpiParams ep = epi(EPI_RESADD, S, d);
            ep.bias = L.co_b; ep.out = c->dx; ep.resid = c->dx;
            produce_ln(ep);
            if (!(skip & 16)) WIPA_TRY(gemm(c, plainA(c->dattn, S, d), L.co_w, S, d, d, ep, c->bn_dec, st));
        }
        }
        if (!lnf && !(skip & 1)) WIPA_TRY(ln(c, c->dx, L.ln3_w, L.ln3_b, c->dh, S, st));
        {
            EpiParams ep = epi(EPI_GELU, S, ffn);
            ep.bias = L.fc1_b; ep.out = c->dffn; ep.out_h16 = c->bf; ep.gelu_fast = c->bf;
            if (lnf) consume_ln(ep, c->fc1_c[l], c->fc1_bf[l]);
            // N = ffn tiles of 32 columns would not fit one wave once S needs two M tiles: use 64-wide tiles then
            if (!(skip & 32)) WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), lnf ? c->fc1_wf[l] : L.fc1_w, S, ffn, d, ep, bn_fc1, st));
        }
        {
            EpiParams ep = epi(EPI_RESADD, S, d);
            ep.bias = L.fc2_b; ep.out = c->dx; ep.resid = c->dx;
            if (c->bf && c->splitk && (!many || c->sk_many)) { ep.sk_part = c->sk_part; ep.sk_count = c->sk_count; ep.sk_splits = wide_sk ? sk_splits : 0; }    // K = ffn is long: split-K
            produce_ln(ep);
            if (!(skip & 64)) WIPA_TRY(gemm(c, plainA(c->dffn, S, ffn), L.fc2_w, S, d, ffn, ep, sk_bn, st));
        }
    }
    int n_tiles = 1;
    if (logits_mode != 0) {
        if (!lnf) WIPA_TRY(ln(c, c->dx, c->dec_ln_w, c->dec_ln_b, c->dh, S, st));
        if (logits_mode == 1 && c->bf) {
            if (!(skip & 128)) WIPA_TRY(logits_argmax(c, S, &n_tiles, st));
        } else {
            float* dst = logits_mode == 2 ? logits_out : c->logits;
            EpiParams ep = epi(EPI_STORE, S, V);
            ep.out = dst; ep.out_h16 = 0; ep.vec_ok = 0;
            ep.ldo = logits_mode == 2 ? ldo : V;
            ep.o_rpb = 1; ep.o_bstride = ep.ldo;                 // row m -> m * ldo
            if (lnf) consume_ln(ep, c->emb_c, c->emb_bf);
            WIPA_TRY(gemm(c, plainA(lnf ? c->dx16 : c->dh, S, d), lnf ? c->emb_wf : c->tok_emb, S, V, d, ep, c->bn_logits ? c->bn_logits : 128, st));
            if (logits_mode == 1)
                WIPA_TRY(launch_row_argmax(dst, S, V, c->mask_always, c->mask_begin, c->d_step, c->pmax, c->pidx, st));
        }
    }
    if (finalize) WIPA_TRY(launch_greedy_finalize(c->pmax, c->pidx, n_tiles, ds, S, st));
    return WIPA_OK;
}

int upload_mask(wipa_ctx* c, uint32_t* dmask, const int32_t* ids, int n, cudaStream_t st) {
    const int words = (c->a.vocab + 31) / 32;
    std::vector<uint32_t> m(words, 0u);
    for (int i = 0; i < n; ++i) {
        WIPA_CHECK(ids[i] >= 0 && ids[i] < c->a.vocab, WIPA_EINVAL, "suppress id %d outside the vocabulary", ids[i]);
        m[ids[i] >> 5] |= 1u << (ids[i] & 31);
    }
    WIPA_CUDA_CHECK(cudaMemcpyAsync(dmask, m.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    WIPA_CUDA_CHECK(cudaStreamSynchronize(st));     // m is a stack-owned pageable buffer
    return WIPA_OK;
}

int decode_setup(wipa_ctx* c, int S, int beams, const std::vector<int32_t>& forced, int n_forced, int max_new, int eot,
                 DecodeState* ds_out, cudaStream_t st) {
    c->cur_beams = beams;
    WIPA_CUDA_CHECK(cudaMemcpyAsync(c->d_forced, forced.data(), forced.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    WIPA_CUDA_CHECK(cudaStreamSynchronize(st));
    DecodeState ds = make_state(c, n_forced, max_new, eot);
    int n = S * c->pages_per_seq;
    if (S * max_new > n) n = S * max_new;
    const int n_counters = c->max_seqs * c->a.heads;
    if (n_counters > n) n = n_counters;
    decode_init_kernel<<<cdiv(n, 256), 256, 0, st>>>(ds, c->block_table, c->utt_of_seq, S, beams, c->pages_per_seq,
                                                      c->ca_counters, n_counters);
    WIPA_LAUNCHED();
    *ds_out = ds;
    return WIPA_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int wipa_ctx_create(const wipa_arch* arch, int max_batch, int max_beams, wipa_ctx** out) {
    WIPA_CHECK(arch && out, WIPA_EINVAL, "wipa_ctx_create: null argument");
    WIPA_CHECK(arch->d_model % 128 == 0 && arch->d_model <= 1280, WIPA_EUNSUPPORTED, "d_model %d must be 128*k <= 1280", arch->d_model);
    WIPA_CHECK(arch->heads * WIPA_HEAD_DIM == arch->d_model, WIPA_EUNSUPPORTED, "head_dim must be 64 (d_model %d, heads %d)",
               arch->d_model, arch->heads);
    WIPA_CHECK(arch->ffn % 64 == 0 && arch->n_mels % 16 == 0 && arch->vocab > 0, WIPA_EINVAL, "bad ffn / n_mels / vocab");
    WIPA_CHECK(arch->dtype == WIPA_DTYPE_F32 || arch->dtype == WIPA_DTYPE_BF16 || arch->dtype == WIPA_DTYPE_F16, WIPA_EINVAL,
               "bad dtype %d", arch->dtype);
    WIPA_CHECK(arch->dtype == WIPA_DTYPE_F32 || arch->dtype == WIPA_H16_DTYPE, WIPA_EUNSUPPORTED,
               "this build of libwipa computes its 16-bit path in " WIPA_H16_NAME " (dtype %d); load %s for dtype %d", WIPA_H16_DTYPE,
               arch->dtype == WIPA_DTYPE_BF16 ? "libwipa_bf16.so" : "libwipa.so", arch->dtype);
    WIPA_CHECK(max_batch >= 1 && max_beams >= 1 && max_batch * max_beams <= 4096, WIPA_EINVAL, "bad max_batch / max_beams");
    int dev = 0, major = 0;
    WIPA_CUDA_CHECK(cudaGetDevice(&dev));
    WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    WIPA_CHECK(major == 10, WIPA_EUNSUPPORTED, "libwipa needs an sm_100a device (found compute capability major %d)", major);

    wipa_ctx* c = new wipa_ctx();
    c->a = *arch;
    c->max_batch = max_batch; c->max_beams = max_beams; c->max_seqs = max_batch * max_beams;
    c->bf = arch->dtype != WIPA_DTYPE_F32;
    c->esz = c->bf ? 2 : 4;
    c->enc_mb = env_int("WIPA_ENC_MB", 32);
    if (c->enc_mb > max_batch) c->enc_mb = max_batch;
    c->pages_per_seq = WIPA_MAX_TGT / WIPA_PAGE;
    c->bn_enc = env_int("WIPA_BN_ENC", 0);
    c->bn_dec = env_int("WIPA_BN_DEC", 32);
    c->bn_logits = env_int("WIPA_BN_LOGITS", 0);
    c->ca_split = env_int("WIPA_CA_SPLIT", cross_attention_default_split((int)c->esz, max_batch, arch->heads));
    c->n_logit_tiles = c->bn_logits == 0 ? 2 * cdiv(arch->vocab, 256) : cdiv(arch->vocab, c->bn_logits);
    c->enc_attn_simt = env_int("WIPA_ENC_ATTN_SIMT", 0);
    c->skip_mask = env_int("WIPA_SKIP_MASK", 0);
    c->splitk = env_int("WIPA_SPLITK", 1);
    c->sk_bn = env_int("WIPA_SK_BN", 32);      // 64 x 6 splits measured slower (2704 vs 2406 us per step): the ticketed reduction, not the A re-reads, is what split-K costs
    c->sk_splits = env_int("WIPA_SK_SPLITS", 6);
    // S > 256 (three or four M tiles): tiles wide enough that every node still fits ONE wave of 148 CTAs
    c->bn_qkv_wide = env_int("WIPA_BN_QKV_WIDE", 64);
    // measured at small / 512 sequences (us per decode step): qkv 64-column tiles 4036 vs 4074 with 32; fc1 64 / 128 and the
    // head-batched q' 128 / 256 within noise; fc2 WITHOUT split-K 3936 vs 4018 - 4036 with any split
    c->bn_fc1_wide = env_int("WIPA_BN_FC1_WIDE", 64);
    c->bn_xlq2_wide = env_int("WIPA_BN_XLQ2_WIDE", 128);
    c->sk_wide = env_int("WIPA_SK_WIDE", 0);
    c->sk_many = env_int("WIPA_SPLITK_MANY", 0);
    c->persistent_min_tiles = env_int("WIPA_PERSISTENT_MIN_TILES", 2 * 148);
    c->bn_xlq = env_int("WIPA_BN_XLQ", 0);
    // latent cross-attention gives every SM whole sequences, so it wants about a wave of them; below that the stream-K
    // kernel over per-layer K / V (exactly balanced at any size) is faster.  WIPA_XATTN_LATENT = 1 / 0 forces either.
    c->xlat = (c->bf && env_int("WIPA_XATTN_LATENT", max_batch * max_beams >= 96 ? 1 : 0) != 0 &&
               (cross_attention_latent_supported(arch->heads) || cross_attention_latent_wide_supported(arch->heads))) ? 1 : 0;
    c->xl_wide = (c->xlat && cross_attention_latent_wide_supported(arch->heads)) ? 1 : 0;
    c->lnf = (c->bf && env_int("WIPA_LN_FOLD", 1) != 0 && arch->d_model % WIPA_LN_PIECE == 0) ? 1 : 0;
    c->xl_q2 = (c->xlat && env_int("WIPA_XL_Q2STEP", 1) != 0) ? 1 : 0;
    c->xl_qfused = (c->xl_q2 && env_int("WIPA_XL_QFUSED", 1) != 0) ? 1 : 0;
    c->xl_o2 = (c->xlat && env_int("WIPA_XL_O2STEP", 1) != 0) ? 1 : 0;
    memset(&c->mel_tables, 0, sizeof(c->mel_tables));

    const int d = arch->d_model, H = arch->heads, ffn = arch->ffn, V = arch->vocab, S = c->max_seqs, mb = c->enc_mb;
    size_t tbytes = 0, fbytes = 0, mbytes = 0;
    layout_weights(c, nullptr, nullptr, nullptr, &tbytes, &fbytes, &mbytes);
    void *tbase = nullptr, *fbase = nullptr, *mbase = nullptr;
    int r = WIPA_OK;
#define CTX_TRY(expr) do { r = (expr); if (r != WIPA_OK) { wipa_ctx_destroy(c); return r; } } while (0)
    CTX_TRY(ctx_alloc(c, &tbase, tbytes, true));
    CTX_TRY(ctx_alloc(c, &fbase, fbytes, true));
    if (mbytes > 0) CTX_TRY(ctx_alloc(c, &mbase, mbytes, true));
    layout_weights(c, (char*)tbase, (char*)fbase, (char*)mbase, &tbytes, &fbytes, &mbytes);
    {
        std::unordered_set<std::string> canon;
        for (auto& kv : c->slots) canon.insert(kv.second.canon);
        c->n_required = canon.size();
    }
    const size_t e = c->esz;
    const size_t T3 = WIPA_N_FRAMES, T = WIPA_T_ENC;
    CTX_TRY(ctx_alloc(c, &c->mel_rows, (size_t)mb * (T3 + 2) * arch->n_mels * e, true));
    CTX_TRY(ctx_alloc(c, &c->conv1_out, (size_t)mb * (T3 + 2) * d * e, true));      // rows 0 / 3001 stay zero
    CTX_TRY(ctx_alloc(c, (void**)&c->ex, (size_t)mb * T * d * 4, false));
    CTX_TRY(ctx_alloc(c, &c->eh, (size_t)mb * T * d * e, false));
    CTX_TRY(ctx_alloc(c, &c->eqkv, (size_t)3 * mb * T * d * e, false));
    CTX_TRY(ctx_alloc(c, &c->eattn, (size_t)mb * T * d * e, false));
    CTX_TRY(ctx_alloc(c, &c->effn, (size_t)mb * T * ffn * e, false));
    CTX_TRY(ctx_alloc(c, &c->enc_T, (size_t)mb * T * d * e, false));
    c->xkv_which_stride = (size_t)max_batch * H * T * 64;
    if (c->xlat) {
        // no per-layer cross-KV: the encoder output itself, plus the folded projections and the [S, H*d] rows around the kernel
        c->xkv = nullptr;
        c->xl_tiled = c->xl_wide ? 1 : env_int("WIPA_XL_TILED", 1);     // the 20-head kernel only reads the tiled layout
        c->xkv_bytes = c->xl_tiled ? (size_t)max_batch * cross_attention_latent_tiled_elems(H, (int)T) * e : (size_t)max_batch * T * d * e;
        CTX_TRY(ctx_alloc(c, &c->enc_lat, c->xkv_bytes, true));          // zeroed: the tiled layout pads every utterance to whole chunks
        CTX_TRY(ctx_alloc(c, &c->dqlat, (size_t)S * H * d * e, false));
        CTX_TRY(ctx_alloc(c, &c->dclat, (size_t)S * H * d * e, false));
        {
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            c->xl_part_floats = c->xl_wide ? cross_attention_latent_wide_scratch_floats(H, S, n_sm) : cross_attention_latent_scratch_floats(H, S, n_sm);
            CTX_TRY(ctx_alloc(c, (void**)&c->xl_part, c->xl_part_floats * 4, false));
        }
        c->xlq_w.assign(arch->dec_layers, nullptr); c->xlo_w.assign(arch->dec_layers, nullptr);
        c->xlq_b.assign(arch->dec_layers, nullptr); c->xlo_b.assign(arch->dec_layers, nullptr);
        c->xl_wkt.assign(arch->dec_layers, nullptr);
        if (c->xl_q2) CTX_TRY(ctx_alloc(c, &c->dq16, (size_t)S * d * e, false));
        for (int l = 0; l < arch->dec_layers; ++l) {
            if (c->xl_q2) CTX_TRY(ctx_alloc(c, &c->xl_wkt[l], (size_t)d * d * e, false));
            else {
                CTX_TRY(ctx_alloc(c, &c->xlq_w[l], (size_t)H * d * d * e, false));
                CTX_TRY(ctx_alloc(c, (void**)&c->xlq_b[l], (size_t)H * d * 4, false));
            }
            if (!c->xl_o2) {
                CTX_TRY(ctx_alloc(c, &c->xlo_w[l], (size_t)H * d * d * e, false));
                CTX_TRY(ctx_alloc(c, (void**)&c->xlo_b[l], (size_t)d * 4, false));
            }
        }
    } else {
        c->xkv_bytes = (size_t)arch->dec_layers * 2 * c->xkv_which_stride * e;
        CTX_TRY(ctx_alloc(c, &c->xkv, c->xkv_bytes, false));
    }
    if (c->lnf) {
        const int Ld = arch->dec_layers;
        CTX_TRY(ctx_alloc(c, &c->dx16, (size_t)S * d * e, false));
        CTX_TRY(ctx_alloc(c, (void**)&c->dstats, (size_t)S * (d / WIPA_LN_PIECE) * 2 * 4, true));
        CTX_TRY(ctx_alloc(c, &c->emb_wf, (size_t)V * d * e, false));
        CTX_TRY(ctx_alloc(c, (void**)&c->emb_c, (size_t)V * 4, false));
        CTX_TRY(ctx_alloc(c, (void**)&c->emb_bf, (size_t)V * 4, false));
        c->qkv_wf.assign(Ld, nullptr); c->cq_wf.assign(Ld, nullptr); c->fc1_wf.assign(Ld, nullptr);
        c->qkv_c.assign(Ld, nullptr); c->qkv_bf.assign(Ld, nullptr); c->cq_c.assign(Ld, nullptr); c->cq_bf.assign(Ld, nullptr);
        c->fc1_c.assign(Ld, nullptr); c->fc1_bf.assign(Ld, nullptr); c->xlq_c.assign(Ld, nullptr); c->xlq_beff.assign(Ld, nullptr);
        for (int l = 0; l < Ld; ++l) {
            CTX_TRY(ctx_alloc(c, &c->qkv_wf[l], (size_t)3 * d * d * e, false));
            CTX_TRY(ctx_alloc(c, (void**)&c->qkv_c[l], (size_t)3 * d * 4, false));
            CTX_TRY(ctx_alloc(c, (void**)&c->qkv_bf[l], (size_t)3 * d * 4, false));
            CTX_TRY(ctx_alloc(c, &c->fc1_wf[l], (size_t)ffn * d * e, false));
            CTX_TRY(ctx_alloc(c, (void**)&c->fc1_c[l], (size_t)ffn * 4, false));
            CTX_TRY(ctx_alloc(c, (void**)&c->fc1_bf[l], (size_t)ffn * 4, false));
            if (c->xlat && !c->xl_q2) {
                CTX_TRY(ctx_alloc(c, (void**)&c->xlq_c[l], (size_t)H * d * 4, false));
                CTX_TRY(ctx_alloc(c, (void**)&c->xlq_beff[l], (size_t)d * 4, false));
            } else {
                CTX_TRY(ctx_alloc(c, &c->cq_wf[l], (size_t)d * d * e, false));
                CTX_TRY(ctx_alloc(c, (void**)&c->cq_c[l], (size_t)d * 4, false));
                CTX_TRY(ctx_alloc(c, (void**)&c->cq_bf[l], (size_t)d * 4, false));
            }
        }
    }
    c->pool_layer_stride = (size_t)S * c->pages_per_seq * H * WIPA_PAGE * 64;
    CTX_TRY(ctx_alloc(c, &c->kpool, (size_t)arch->dec_layers * c->pool_layer_stride * e, false));
    CTX_TRY(ctx_alloc(c, &c->vpool, (size_t)arch->dec_layers * c->pool_layer_stride * e, false));
    CTX_TRY(ctx_alloc(c, (void**)&c->dx, (size_t)S * d * 4, false));
    CTX_TRY(ctx_alloc(c, (void**)&c->dq, (size_t)S * d * 4, false));
    CTX_TRY(ctx_alloc(c, &c->dh, (size_t)S * d * e, false));
    CTX_TRY(ctx_alloc(c, &c->dattn, (size_t)S * d * e, false));
    CTX_TRY(ctx_alloc(c, &c->dffn, (size_t)S * ffn * e, false));
    CTX_TRY(ctx_alloc(c, (void**)&c->ca_part, (size_t)S * H * 64 * 66 * 4, false));
    {
        const size_t tiles = (size_t)cdiv(d, 32) * cdiv(S, 64);        // N tiles of 32 columns x M tiles (64-row tiles at S <= 64)
        // a partial tile is 128 rows x BN fp32 per (n-tile, m-tile, split): d * 512 bytes per (m-tile, split) whatever BN is
        const size_t sk_bytes = (size_t)d * 512 * cdiv(S, 64) * (size_t)(c->sk_splits > 3 ? c->sk_splits : 3);
        CTX_TRY(ctx_alloc(c, (void**)&c->sk_part, sk_bytes > tiles * 3 * 128 * 64 * 4 ? sk_bytes : tiles * 3 * 128 * 64 * 4, false));
        CTX_TRY(ctx_alloc(c, (void**)&c->sk_count, tiles * 4, true));
    }
    CTX_TRY(ctx_alloc(c, (void**)&c->ca_counters, (size_t)S * H * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->pmax, (size_t)S * c->n_logit_tiles * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->pidx, (size_t)S * c->n_logit_tiles * 4, true));
    if (!c->bf || max_beams > 1) CTX_TRY(ctx_alloc(c, (void**)&c->logits, (size_t)S * V * 4, false));
    else c->logits = nullptr;
    if (max_beams > 1) {
        const size_t SL = (size_t)S * WIPA_MAX_TGT;
        CTX_TRY(ctx_alloc(c, (void**)&c->b_flip, 256, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_run_seq, 2 * SL * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_fin_seq, 2 * SL * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_anc, 2 * SL * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_run_score, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_run_score_next, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_fin_score, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_fin_score_next, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_fin_done, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_fin_done_next, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_unsat, (size_t)S * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_cand_val, (size_t)S * 16 * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_cand_idx, (size_t)S * 16 * 4, true));
        CTX_TRY(ctx_alloc(c, (void**)&c->b_prompt, (size_t)WIPA_MAX_TGT * 4, true));
    }
    CTX_TRY(ctx_alloc(c, (void**)&c->block_table, (size_t)S * c->pages_per_seq * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->utt_of_seq, (size_t)S * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->d_pos, 256, true));
    c->d_step = c->d_pos + 1; c->d_n_done = c->d_pos + 2;
    CTX_TRY(ctx_alloc(c, (void**)&c->d_cur_tok, (size_t)S * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->d_done, (size_t)S * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->d_forced, (size_t)S * (WIPA_MAX_TGT + 1) * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->d_out_ids, (size_t)S * WIPA_MAX_TGT * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->d_out_len, (size_t)S * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->mask_always, (size_t)((V + 31) / 32) * 4, true));
    CTX_TRY(ctx_alloc(c, (void**)&c->mask_begin, (size_t)((V + 31) / 32) * 4, true));
    if (cudaMallocHost((void**)&c->h_pinned, 256) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        wipa_set_error("cudaMallocHost / cudaStreamCreate failed");
        cudaGetLastError();
        wipa_ctx_destroy(c);
        return WIPA_ENOMEM;
    }
    {
        // The encoder's position table is not a learned weight: Whisper fixes it to sinusoids (HF:models/whisper/
        // modeling_whisper.py:55-65), and MLX-format checkpoints do not carry it at all (mlx_whisper keeps it as the private
        // `_positional_embedding`).  Pre-fill it, so the slot is optional; a state dict that has the tensor overwrites it.
        const int half = d / 2;
        std::vector<float> pos((size_t)T * d);
        const double inc = log(10000.0) / (double)(half - 1);
        for (size_t t = 0; t < T; ++t)
            for (int i = 0; i < half; ++i) {
                const double a = (double)t * exp(-inc * (double)i);
                pos[t * d + i] = (float)sin(a);
                pos[t * d + half + i] = (float)cos(a);
            }
        if (cudaMemcpy(c->enc_pos, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            wipa_set_error("cudaMemcpy of the encoder position table failed");
            cudaGetLastError();
            wipa_ctx_destroy(c);
            return WIPA_ECUDA;
        }
        c->loaded.insert("model.encoder.embed_positions.weight");
    }
#undef CTX_TRY
    *out = c;
    return WIPA_OK;
}

extern "C" int wipa_ctx_destroy(wipa_ctx* c) {
    if (!c) return WIPA_OK;
    cudaDeviceSynchronize();
    for (auto& g : c->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    for (void* p : c->allocs) cudaFree(p);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    delete c;
    return WIPA_OK;
}

extern "C" int wipa_ctx_load_weights(wipa_ctx* c, const wipa_tensor_desc* tensors, int n, void* stream) {
    WIPA_CHECK(c && (tensors || n == 0), WIPA_EINVAL, "wipa_ctx_load_weights: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n; ++i) {
        const wipa_tensor_desc& t = tensors[i];
        WIPA_CHECK(t.name && t.data, WIPA_EINVAL, "tensor %d: null name or data", i);
        auto it = c->slots.find(t.name);
        WIPA_CHECK(it != c->slots.end(), WIPA_EINVAL, "unknown tensor name '%s'", t.name);
        const Slot& s = it->second;
        WIPA_CHECK(t.numel == s.numel, WIPA_EINVAL, "tensor '%s': %lld elements, expected %lld", t.name, (long long)t.numel,
                   (long long)s.numel);
        if (s.kind == SLOT_CONV) WIPA_TRY(launch_conv_weight(t.data, s.dst, s.N, s.C, c->bf, st));
        else WIPA_TRY(launch_convert(t.data, s.dst, s.numel, s.scale, s.kind == SLOT_T ? (int)c->bf : 0, st));
        if (s.master != nullptr) WIPA_TRY(launch_convert(t.data, s.master, s.numel, s.scale, 0, st));
        c->loaded.insert(s.canon);
    }
    // the weights changed: graphs stay valid (same buffers), cached cross-KV and folded projections do not
    c->n_utts = 0;
    c->xlat_ready = false;
    c->lnf_ready = false;
    return WIPA_OK;
}

extern "C" int wipa_logmel(const float* audio, int B, int n_mels, float* mel, void* stream) {
    WIPA_CHECK(B >= 0 && (n_mels == 80 || n_mels == 128), WIPA_EINVAL, "wipa_logmel: B=%d n_mels=%d (80 or 128)", B, n_mels);
    if (B == 0) return WIPA_OK;
    WIPA_CHECK(audio && mel, WIPA_EINVAL, "wipa_logmel: null pointer");
    // constant tables (windowed DFT matrix, banded filterbank) are built once per (device, n_mels)
    struct Cached { LogmelTables t; float* clipmax; int cap; bool ok; };
    static Cached cache[16][2];
    int dev = 0;
    WIPA_CUDA_CHECK(cudaGetDevice(&dev));
    WIPA_CHECK(dev < 16, WIPA_EUNSUPPORTED, "device index %d", dev);
    Cached& cc = cache[dev][n_mels == 128];
    if (!cc.ok) {
        WIPA_TRY(logmel_tables_create(n_mels, &cc.t));
        cc.clipmax = nullptr; cc.cap = 0; cc.ok = true;
    }
    if (cc.cap < B) {
        if (cc.clipmax) { WIPA_CUDA_CHECK(cudaDeviceSynchronize()); cudaFree(cc.clipmax); }
        WIPA_CUDA_CHECK(cudaMalloc(&cc.clipmax, sizeof(float) * (size_t)B));
        cc.cap = B;
    }
    return launch_logmel(cc.t, audio, B, mel, cc.clipmax, (cudaStream_t)stream);
}

extern "C" int wipa_encode(wipa_ctx* c, const float* mel, int B, float* enc_out, void* stream) {
    WIPA_CHECK(c && mel, WIPA_EINVAL, "wipa_encode: null argument");
    WIPA_CHECK(B >= 1 && B <= c->max_batch, WIPA_EINVAL, "wipa_encode: B=%d outside 1..%d", B, c->max_batch);
    WIPA_TRY(require_weights(c));
    cudaStream_t st = (cudaStream_t)stream;
    WIPA_TRY(xlat_prepare(c, st));
    WIPA_TRY(lnf_prepare(c, st));
    const size_t mel_clip = (size_t)c->a.n_mels * WIPA_N_FRAMES, enc_clip = (size_t)WIPA_T_ENC * c->a.d_model;
    for (int u0 = 0; u0 < B; u0 += c->enc_mb) {
        const int nb = (B - u0) < c->enc_mb ? (B - u0) : c->enc_mb;
        WIPA_TRY(encode_chunk(c, mel + u0 * mel_clip, u0, nb, enc_out ? enc_out + u0 * enc_clip : nullptr, st));
    }
    c->n_utts = B;
    c->step_pos = 0;
    return WIPA_OK;
}

extern "C" int wipa_set_audio_features(wipa_ctx* c, const float* enc_out, int B, void* stream) {
    WIPA_CHECK(c && enc_out, WIPA_EINVAL, "wipa_set_audio_features: null argument");
    WIPA_CHECK(B >= 1 && B <= c->max_batch, WIPA_EINVAL, "wipa_set_audio_features: B=%d outside 1..%d", B, c->max_batch);
    WIPA_TRY(require_weights(c));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t enc_clip = (size_t)WIPA_T_ENC * c->a.d_model;
    WIPA_TRY(xlat_prepare(c, st));
    WIPA_TRY(lnf_prepare(c, st));
    if (c->xlat && c->xl_tiled) WIPA_TRY(launch_lat_tile(enc_out, 0, (h16*)c->enc_lat, B, WIPA_T_ENC, c->a.heads, cross_attention_latent_keys(c->a.heads), st));
    else if (c->xlat) WIPA_TRY(launch_convert(enc_out, c->enc_lat, (long long)B * enc_clip, 1.0f, 1, st));
    else for (int u0 = 0; u0 < B; u0 += c->enc_mb) {
        const int nb = (B - u0) < c->enc_mb ? (B - u0) : c->enc_mb;
        WIPA_TRY(launch_convert(enc_out + u0 * enc_clip, c->enc_T, (long long)nb * enc_clip, 1.0f, c->bf, st));
        WIPA_TRY(cross_kv_project(c, u0, nb, st));
    }
    c->n_utts = B;
    c->step_pos = 0;
    return WIPA_OK;
}

static int check_decode_args(wipa_ctx* c, int B, const wipa_decode_opts* o, int32_t* out_ids, int32_t* out_len) {
    WIPA_CHECK(c && o && out_ids && out_len, WIPA_EINVAL, "decode: null argument");
    WIPA_CHECK(c->n_utts > 0, WIPA_ESTATE, "decode before wipa_encode / wipa_set_audio_features");
    WIPA_CHECK(B == c->n_utts, WIPA_EINVAL, "decode: B=%d but %d utterances are encoded", B, c->n_utts);
    WIPA_CHECK(o->prompt && o->prompt_len >= 1, WIPA_EINVAL, "decode: empty prompt");
    WIPA_CHECK(o->max_new >= 1 && o->prompt_len + o->max_new <= WIPA_MAX_TGT, WIPA_EINVAL,
               "decode: prompt_len %d + max_new %d exceeds %d target positions", o->prompt_len, o->max_new, WIPA_MAX_TGT);
    WIPA_CHECK(o->eot >= 0 && o->eot < c->a.vocab, WIPA_EINVAL, "decode: eot %d outside the vocabulary", o->eot);
    WIPA_CHECK((o->n_suppress == 0 || o->suppress) && (o->n_begin_suppress == 0 || o->begin_suppress), WIPA_EINVAL,
               "decode: suppress list pointer missing");
    for (int i = 0; i < o->prompt_len; ++i)
        WIPA_CHECK(o->prompt[i] >= 0 && o->prompt[i] < c->a.vocab, WIPA_EINVAL, "decode: prompt id %d outside the vocabulary", o->prompt[i]);
    return WIPA_OK;
}

extern "C" int wipa_decode_greedy(wipa_ctx* c, int B, const wipa_decode_opts* o, int32_t* out_ids, int32_t* out_len, void* stream) {
    WIPA_TRY(check_decode_args(c, B, o, out_ids, out_len));
    WIPA_TRY(require_weights(c));
    c->step_pos = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = B, P = o->prompt_len, max_new = o->max_new;
    WIPA_TRY(upload_mask(c, c->mask_always, o->suppress, o->n_suppress, st));
    WIPA_TRY(upload_mask(c, c->mask_begin, o->begin_suppress, o->n_begin_suppress, st));
    std::vector<int32_t> forced((size_t)S * P);
    for (int b = 0; b < S; ++b) memcpy(&forced[(size_t)b * P], o->prompt, sizeof(int32_t) * P);
    DecodeState ds;
    WIPA_TRY(decode_setup(c, S, 1, forced, P, max_new, o->eot, &ds, st));

    const int total_steps = P - 1 + max_new;           // P-1 forced positions, then max_new sampled tokens
    int s = 0;
    for (; s < P - 1; ++s) WIPA_TRY(decode_step(c, S, ds, 0, nullptr, 0, st));
    // first sampled step eagerly (also settles every lazily-set function attribute), the rest as graph replays
    WIPA_TRY(decode_step(c, S, ds, 1, nullptr, 0, st));
    ++s;
    const bool use_graph = env_int("WIPA_NO_GRAPH", 0) == 0 && total_steps - s >= 2;
    GraphEntry* ge = nullptr;
    if (use_graph) {
        ge = &c->graphs[{0LL, (long long)S, (long long)P, (long long)max_new, (long long)o->eot, 1LL, 0LL}];
        if (ge->exec == nullptr) {
            cudaGraph_t g = nullptr;
            const int64_t before = g_wipa_launches;
            WIPA_CUDA_CHECK(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int r = decode_step(c, S, ds, 1, nullptr, 0, c->cap_stream);
            cudaError_t e = cudaStreamEndCapture(c->cap_stream, &g);
            if (r != WIPA_OK) { if (g) cudaGraphDestroy(g); return r; }
            WIPA_CUDA_CHECK(e);
            ge->nodes = (int)(g_wipa_launches - before);
            g_wipa_launches = before;                    // capture launched nothing
            WIPA_CUDA_CHECK(cudaGraphInstantiate(&ge->exec, g, 0));
            cudaGraphDestroy(g);
        }
    }
    int* h_done = c->h_pinned;
    *h_done = 0;
    while (s < total_steps) {
        int burst = total_steps - s < 32 ? total_steps - s : 32;
        for (int i = 0; i < burst; ++i) {
            if (ge) { WIPA_CUDA_CHECK(cudaGraphLaunch(ge->exec, st)); g_wipa_launches += ge->nodes; }
            else WIPA_TRY(decode_step(c, S, ds, 1, nullptr, 0, st));
        }
        s += burst;
        if (s < total_steps) {       // all rows finished? (HF stops as soon as every row has emitted EOS)
            WIPA_CUDA_CHECK(cudaMemcpyAsync(h_done, c->d_n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
            WIPA_CUDA_CHECK(cudaStreamSynchronize(st));
            if (*h_done >= S) break;
        }
    }
    c->decode_steps += s;
    WIPA_CUDA_CHECK(cudaMemcpyAsync(out_ids, c->d_out_ids, sizeof(int32_t) * (size_t)S * max_new, cudaMemcpyDeviceToDevice, st));
    WIPA_CUDA_CHECK(cudaMemcpyAsync(out_len, c->d_out_len, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToDevice, st));
    return WIPA_OK;
}

extern "C" int wipa_decode_logits(wipa_ctx* c, int B, const int32_t* tokens, int T, float* logits, void* stream) {
    WIPA_CHECK(c && tokens && logits, WIPA_EINVAL, "wipa_decode_logits: null argument");
    WIPA_CHECK(c->n_utts > 0 && B == c->n_utts, WIPA_ESTATE, "wipa_decode_logits: %d utterances encoded, B=%d", c->n_utts, B);
    WIPA_CHECK(T >= 1 && T <= WIPA_MAX_TGT, WIPA_EINVAL, "wipa_decode_logits: T=%d", T);
    WIPA_TRY(require_weights(c));
    c->step_pos = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int V = c->a.vocab;
    // T forced tokens + one dummy so that every step stays on the teacher-forced branch of the finalize kernel
    std::vector<int32_t> forced((size_t)B * (T + 1), 0);
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) {
            const int32_t id = tokens[(size_t)b * T + t];
            WIPA_CHECK(id >= 0 && id < V, WIPA_EINVAL, "wipa_decode_logits: token %d outside the vocabulary", id);
            forced[(size_t)b * (T + 1) + t] = id;
        }
    DecodeState ds;
    WIPA_TRY(decode_setup(c, B, 1, forced, T + 1, 0, 0, &ds, st));
    for (int t = 0; t < T; ++t)
        WIPA_TRY(decode_step(c, B, ds, 2, logits + (size_t)t * V, (long long)T * V, st));
    c->decode_steps += T;
    return WIPA_OK;
}

// ---- stepwise decoding with the logits handed back to the caller (host-side logits processors, sampling, fallback) ----
__global__ void decode_advance_kernel(int* pos, int* step) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) { *pos += 1; *step += 1; }
}

__global__ void decode_set_tokens_kernel(const int32_t* __restrict__ tokens, int* __restrict__ cur_tok, int n, int vocab) {
    pdl_wait();
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int t = tokens[i]; cur_tok[i] = (t >= 0 && t < vocab) ? t : 0; }
}

extern "C" int wipa_decode_begin(wipa_ctx* c, int B, const int32_t* tokens, int T, float* logits, void* stream) {
    WIPA_CHECK(c && tokens && logits, WIPA_EINVAL, "wipa_decode_begin: null argument");
    WIPA_CHECK(c->n_utts > 0 && B == c->n_utts, WIPA_ESTATE, "wipa_decode_begin: %d utterances encoded, B=%d", c->n_utts, B);
    WIPA_CHECK(T >= 1 && T <= WIPA_MAX_TGT, WIPA_EINVAL, "wipa_decode_begin: T=%d", T);
    WIPA_TRY(require_weights(c));
    cudaStream_t st = (cudaStream_t)stream;
    const int V = c->a.vocab;
    std::vector<int32_t> forced((size_t)B * (T + 1), 0);        // + one dummy: every position stays on the teacher-forced branch
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) {
            const int32_t id = tokens[(size_t)b * T + t];
            WIPA_CHECK(id >= 0 && id < V, WIPA_EINVAL, "wipa_decode_begin: token %d outside the vocabulary", id);
            forced[(size_t)b * (T + 1) + t] = id;
        }
    DecodeState ds;
    WIPA_TRY(decode_setup(c, B, 1, forced, T + 1, 0, 0, &ds, st));
    for (int t = 0; t < T - 1; ++t) WIPA_TRY(decode_step(c, B, ds, 0, nullptr, 0, st));
    WIPA_TRY(decode_step(c, B, ds, 2, logits, (long long)V, st, /*beam=*/false, /*finalize=*/false));
    decode_advance_kernel<<<1, 32, 0, st>>>(c->d_pos, c->d_step);
    WIPA_LAUNCHED();
    c->decode_steps += T;
    c->step_pos = T;
    return WIPA_OK;
}

extern "C" int wipa_decode_next(wipa_ctx* c, int B, const int32_t* tokens_dev, float* logits, void* stream) {
    WIPA_CHECK(c && tokens_dev && logits, WIPA_EINVAL, "wipa_decode_next: null argument");
    WIPA_CHECK(c->step_pos > 0 && B == c->n_utts, WIPA_ESTATE, "wipa_decode_next before wipa_decode_begin (or B=%d != %d)", B, c->n_utts);
    WIPA_CHECK(c->step_pos < WIPA_MAX_TGT, WIPA_EINVAL, "wipa_decode_next: all %d target positions are used", WIPA_MAX_TGT);
    cudaStream_t st = (cudaStream_t)stream;
    decode_set_tokens_kernel<<<cdiv(B, 256), 256, 0, st>>>(tokens_dev, c->d_cur_tok, B, c->a.vocab);
    WIPA_LAUNCHED();
    DecodeState ds = make_state(c, 1, 0, 0);
    WIPA_TRY(decode_step(c, B, ds, 2, logits, (long long)c->a.vocab, st, /*beam=*/false, /*finalize=*/false));
    decode_advance_kernel<<<1, 32, 0, st>>>(c->d_pos, c->d_step);
    WIPA_LAUNCHED();
    c->decode_steps += 1;
    c->step_pos += 1;
    return WIPA_OK;
}

extern "C" int wipa_decode_beam(wipa_ctx* c, int B, int beams, float length_penalty, const wipa_decode_opts* o,
                                int32_t* out_ids, int32_t* out_len, void* stream) {
    WIPA_TRY(check_decode_args(c, B, o, out_ids, out_len));
    WIPA_TRY(require_weights(c));
    WIPA_CHECK(beams >= 2 && beams <= 8, WIPA_EINVAL, "wipa_decode_beam: beams=%d outside 2..8 (use wipa_decode_greedy for 1)", beams);
    WIPA_CHECK(beams <= c->max_beams && B * beams <= c->max_seqs, WIPA_EINVAL,
               "wipa_decode_beam: B=%d x beams=%d exceeds the context (max_batch %d, max_beams %d)", B, beams, c->max_batch, c->max_beams);
    c->step_pos = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = B * beams, P = o->prompt_len, max_new = o->max_new, V = c->a.vocab, keep = 2 * beams;
    const int L = P + max_new;                                 // HF's max_length
    WIPA_TRY(upload_mask(c, c->mask_always, o->suppress, o->n_suppress, st));
    WIPA_TRY(upload_mask(c, c->mask_begin, o->begin_suppress, o->n_begin_suppress, st));
    std::vector<int32_t> forced((size_t)S * P);
    for (int b = 0; b < S; ++b) memcpy(&forced[(size_t)b * P], o->prompt, sizeof(int32_t) * P);
    WIPA_CUDA_CHECK(cudaMemcpyAsync(c->b_prompt, o->prompt, sizeof(int32_t) * P, cudaMemcpyHostToDevice, st));
    DecodeState ds;
    WIPA_TRY(decode_setup(c, S, beams, forced, P, max_new, o->eot, &ds, st));   // utt_of_seq[s] = s / beams: beams share the cross-KV
    c->beam_L = L;
    BeamState bs;
    bs.pos = c->d_pos; bs.step = c->d_step; bs.flip = c->b_flip; bs.cur_tok = c->d_cur_tok;
    bs.run_seq = c->b_run_seq; bs.fin_seq = c->b_fin_seq; bs.anc = c->b_anc;
    bs.run_score = c->b_run_score; bs.run_score_next = c->b_run_score_next;
    bs.fin_score = c->b_fin_score; bs.fin_score_next = c->b_fin_score_next;
    bs.fin_done = c->b_fin_done; bs.fin_done_next = c->b_fin_done_next;
    bs.unsat = c->b_unsat; bs.n_done = c->d_n_done;
    bs.beams = beams; bs.L = L; bs.prompt_len = P; bs.max_length = L; bs.eot = o->eot; bs.length_penalty = length_penalty;
    WIPA_TRY(launch_beam_init(bs, c->b_prompt, B, st));

    auto beam_step = [&](cudaStream_t s2) -> int {
        // decoder forward for all S running beams with fp32 logits, then the two bookkeeping kernels
        WIPA_TRY(decode_step(c, S, ds, 2, c->logits, (long long)V, s2, /*beam=*/true, /*finalize=*/false));
        WIPA_TRY(launch_beam_row_topk(c->logits, (long long)V, V, c->mask_always, c->mask_begin, c->d_step, c->b_run_score, keep,
                                      c->b_cand_val, c->b_cand_idx, S, s2));
        WIPA_TRY(launch_beam_update(bs, c->b_cand_val, c->b_cand_idx, V, B, s2));
        return WIPA_OK;
    };
    // the P-1 teacher-forced prompt positions: every beam of an utterance consumes the same tokens
    for (int s = 0; s < P - 1; ++s) WIPA_TRY(decode_step(c, S, ds, 0, nullptr, 0, st, /*beam=*/true));
    WIPA_TRY(beam_step(st));                                    // first sampled step eagerly, the rest as graph replays
    int done_steps = 1;
    GraphEntry* ge = nullptr;
    if (env_int("WIPA_NO_GRAPH", 0) == 0 && max_new >= 3) {
        uint32_t lp_bits;
        memcpy(&lp_bits, &length_penalty, 4);
        ge = &c->graphs[{1LL, (long long)S, (long long)P, (long long)max_new, (long long)o->eot, (long long)beams, (long long)lp_bits}];
        if (ge->exec == nullptr) {
            cudaGraph_t g = nullptr;
            const int64_t before = g_wipa_launches;
            WIPA_CUDA_CHECK(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int r = beam_step(c->cap_stream);
            cudaError_t e = cudaStreamEndCapture(c->cap_stream, &g);
            if (r != WIPA_OK) { if (g) cudaGraphDestroy(g); return r; }
            WIPA_CUDA_CHECK(e);
            ge->nodes = (int)(g_wipa_launches - before);
            g_wipa_launches = before;
            WIPA_CUDA_CHECK(cudaGraphInstantiate(&ge->exec, g, 0));
            cudaGraphDestroy(g);
        }
    }
    int* h_done = c->h_pinned;
    *h_done = 0;
    while (done_steps < max_new) {
        const int burst = max_new - done_steps < 16 ? max_new - done_steps : 16;
        for (int i = 0; i < burst; ++i) {
            if (ge) { WIPA_CUDA_CHECK(cudaGraphLaunch(ge->exec, st)); g_wipa_launches += ge->nodes; }
            else WIPA_TRY(beam_step(st));
        }
        done_steps += burst;
        if (done_steps < max_new) {       // HF stops once no utterance can improve its finished beams any more
            WIPA_CUDA_CHECK(cudaMemcpyAsync(h_done, c->d_n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
            WIPA_CUDA_CHECK(cudaStreamSynchronize(st));
            if (*h_done >= B) break;
        }
    }
    c->decode_steps += P - 1 + done_steps;
    return launch_beam_finish(bs, B, max_new, out_ids, out_len, st);
}

extern "C" int wipa_h16_dtype(void) { return WIPA_H16_DTYPE; }

extern "C" int wipa_ctx_get_info(wipa_ctx* c, int what, int64_t* out) {
    WIPA_CHECK(c && out, WIPA_EINVAL, "wipa_ctx_get_info: null argument");
    switch (what) {
        case WIPA_INFO_WORKSPACE_BYTES: *out = (int64_t)c->workspace_bytes; return WIPA_OK;
        case WIPA_INFO_CROSSKV_BYTES: *out = (int64_t)c->xkv_bytes; return WIPA_OK;
        case WIPA_INFO_DECODE_STEPS: *out = c->decode_steps; return WIPA_OK;
        case WIPA_INFO_XATTN_LATENT: *out = c->xlat; return WIPA_OK;
        default: break;
    }
    wipa_set_error("wipa_ctx_get_info: unknown selector %d", what);
    return WIPA_EINVAL;
}

// ---- standalone kernel entry points (tests / roofline) ----------------------------------------------
extern "C" int wipa_test_gemm_h16(const void* A, const void* W, const float* bias, float* C, int M, int N, int K, int block_n,
                                   void* stream) {
    EpiParams ep = epi(EPI_STORE, M, N);
    ep.bias = bias; ep.out = C; ep.out_h16 = 0;
    return launch_gemm_h16(plainA(A, M, K), (const h16*)W, M, N, K, ep, block_n, (cudaStream_t)stream);
}

extern "C" int wipa_test_gemm_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, void* stream) {
    EpiParams ep = epi(EPI_STORE, M, N);
    ep.bias = bias; ep.out = C; ep.out_h16 = 0;
    return launch_gemm_f32(plainA(A, M, K), W, M, N, K, ep, (cudaStream_t)stream);
}

// Epilogue variants of the h16 GEMMs on encoder-shaped problems: M = rows_per_batch * n_batch rows, processed as the
// encoder does (tiles never straddle a batch).  mode: 0 bias, 1 bias + GELU, 2 bias + residual (fp32 in place of `out`),
// 3 bias + q|k|v head split into [3][n_batch][H][rows_per_batch][64] (N = 3 * H * 64).  `out` is h16 when out_h16 != 0.
extern "C" int wipa_test_gemm_epilogue(const void* A, const void* W, const float* bias, const float* resid, void* out,
                                       int rows_per_batch, int n_batch, int N, int K, int mode, int out_h16, int block_n,
                                       void* stream) {
    WIPA_CHECK(A && W && out && rows_per_batch > 0 && n_batch > 0, WIPA_EINVAL, "wipa_test_gemm_epilogue: bad argument");
    WIPA_CHECK(mode >= 0 && mode <= 3, WIPA_EINVAL, "wipa_test_gemm_epilogue: mode must be 0..3");
    const int M = rows_per_batch * n_batch;
    static const int modes[4] = {EPI_STORE, EPI_GELU, EPI_RESADD, EPI_HEADS};
    EpiParams ep = epi(modes[mode], M, N);
    ep.bias = bias; ep.out = out; ep.out_h16 = out_h16 ? 1 : 0;
    ep.o_rpb = rows_per_batch; ep.o_bstride = (long long)rows_per_batch * N;
    if (mode == 1) ep.gelu_fast = ep.out_h16;
    if (mode == 2) {
        WIPA_CHECK(resid != nullptr && !out_h16, WIPA_EINVAL, "wipa_test_gemm_epilogue: residual mode is fp32 and needs resid");
        ep.resid = resid;
    }
    if (mode == 3) {
        WIPA_CHECK(N % (3 * WIPA_HEAD_DIM) == 0, WIPA_EINVAL, "wipa_test_gemm_epilogue: N must be 3 * H * 64");
        ep.d = N / 3; ep.H = ep.d / WIPA_HEAD_DIM; ep.T = rows_per_batch;
        ep.which_stride = (long long)n_batch * ep.H * rows_per_batch * WIPA_HEAD_DIM;
    }
    AOperand a; a.ptr = A; a.lda = K; a.a_rpb = rows_per_batch; a.a_bstride = (long long)rows_per_batch * K; a.n_batch = n_batch;
    return launch_gemm_h16(a, (const h16*)W, M, N, K, ep, block_n, (cudaStream_t)stream);
}

// conv-as-GEMM addressing check: row m = (batch, t) of A starts at A + batch*bstride + t*lda and spans K >= lda elements
extern "C" int wipa_test_gemm_rows(const void* A, int is_h16, long long lda, int rows_per_batch, long long bstride, int n_batch,
                                   const void* W, float* C, int N, int K, int block_n, void* stream) {
    const int M = rows_per_batch * n_batch;
    EpiParams ep = epi(EPI_STORE, M, N);
    ep.out = C; ep.out_h16 = 0;
    AOperand a; a.ptr = A; a.lda = lda; a.a_rpb = rows_per_batch; a.a_bstride = bstride; a.n_batch = n_batch;
    if (is_h16) return launch_gemm_h16(a, (const h16*)W, M, N, K, ep, block_n, (cudaStream_t)stream);
    return launch_gemm_f32(a, (const float*)W, M, N, K, ep, (cudaStream_t)stream);
}

extern "C" int wipa_test_cross_attn(wipa_ctx* c, int B, int layer, const float* q, float* out, void* stream) {
    WIPA_CHECK(c && q, WIPA_EINVAL, "wipa_test_cross_attn: null argument");
    WIPA_CHECK(B >= 1 && B <= c->max_seqs && layer >= 0 && layer < c->a.dec_layers, WIPA_EINVAL, "wipa_test_cross_attn: bad B / layer");
    WIPA_CHECK(!c->xlat, WIPA_ESTATE, "wipa_test_cross_attn: this context keeps no cross-KV (latent cross-attention)");
    cudaStream_t st = (cudaStream_t)stream;
    const char* xk = (const char*)c->xkv + (size_t)(2 * layer) * c->xkv_which_stride * c->esz;
    const char* xv = (const char*)c->xkv + (size_t)(2 * layer + 1) * c->xkv_which_stride * c->esz;
    if (c->bf) WIPA_TRY(launch_cross_attention<h16>(q, (const h16*)xk, (const h16*)xv, nullptr, (h16*)c->dattn, c->ca_part,
                                                     c->ca_counters, B, c->a.heads, c->ca_split, 0, st));
    else WIPA_TRY(launch_cross_attention<float>(q, (const float*)xk, (const float*)xv, nullptr, (float*)c->dattn, c->ca_part,
                                                c->ca_counters, B, c->a.heads, c->ca_split, 0, st));
    if (out) WIPA_TRY(launch_to_f32(c->dattn, c->bf, out, (long long)B * c->a.d_model, st));
    return WIPA_OK;
}

// the fused vocabulary projection + argmax node of a decode step alone, on the S rows currently in the context's
// LayerNorm-output buffer (16-bit contexts only): bench.py times it for its achieved HBM GB/s
extern "C" int wipa_test_logits_argmax(wipa_ctx* c, int S, void* stream) {
    WIPA_CHECK(c && c->bf && S >= 1 && S <= c->max_seqs, WIPA_EINVAL, "wipa_test_logits_argmax: 16-bit context and 1 <= S <= %d", c ? c->max_seqs : 0);
    WIPA_TRY(require_weights(c));
    int n_tiles = 0;
    return logits_argmax(c, S, &n_tiles, (cudaStream_t)stream);
}

// latent cross-attention kernel alone: Qp h16 [S, H, 64H] absorbed queries, E h16 [U, T, 64H], utt_of_seq int32 [S]
// -> C h16 [S, H, 64H] = softmax_t(Qp[s, h] . E[u, t]) E[u].  All device pointers.
extern "C" int wipa_test_cross_attn_latent(const void* Qp, const void* E, int U, const int* utt_of_seq, void* C, int S, int H, int T,
                                           int layout, int beams, void* stream) {
    // 16 heads have both kernels: the row-major layout goes to attn_lat.cu, the tiled ones to attn_lat_wide.cu like in a context
    const bool wide = cross_attention_latent_wide_supported(H) != 0 && (layout != 0 || !cross_attention_latent_supported(H));
    WIPA_CHECK(Qp && E && utt_of_seq && C && S >= 1 && (wide || cross_attention_latent_supported(H)), WIPA_EINVAL,
               "wipa_test_cross_attn_latent: bad argument");
    WIPA_CHECK(!wide || layout != 0, WIPA_EUNSUPPORTED, "wipa_test_cross_attn_latent: 20 heads read the chunk-tiled layout only (layout 1 or 2)");
    WIPA_CHECK(layout >= 0 && layout <= 2, WIPA_EINVAL, "wipa_test_cross_attn_latent: layout 0 (row-major, TMA boxes), 1 (row-major in, "
               "tiled internally) or 2 (already chunk-tiled)");
    // scratch of the standalone entry point: grow-only, per process (contexts own theirs)
    static float* part = nullptr;
    static int* counters = nullptr;
    static h16* tiled = nullptr;
    static size_t part_floats = 0, tiled_elems = 0;
    static int counters_cap = 0;
    int dev = 0, n_sm = 148;
    WIPA_CUDA_CHECK(cudaGetDevice(&dev));
    WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const size_t need = wide ? cross_attention_latent_wide_scratch_floats(H, S, n_sm) : cross_attention_latent_scratch_floats(H, S, n_sm);
    if (need > part_floats) {
        WIPA_CUDA_CHECK(cudaDeviceSynchronize());
        if (part) cudaFree(part);
        WIPA_CUDA_CHECK(cudaMalloc(&part, need * 4));
        part_floats = need;
    }
    if (2 * S > counters_cap) {                                  // 20 heads: one counter per (sequence, head group)
        WIPA_CUDA_CHECK(cudaDeviceSynchronize());
        if (counters) cudaFree(counters);
        WIPA_CUDA_CHECK(cudaMalloc(&counters, (size_t)S * 8));
        WIPA_CUDA_CHECK(cudaMemset(counters, 0, (size_t)S * 8));
        counters_cap = 2 * S;
    }
    const h16* Ein = (const h16*)E;
    if (layout == 1) {
        const size_t elems = (size_t)U * cross_attention_latent_tiled_elems(H, T);
        if (elems > tiled_elems) {
            WIPA_CUDA_CHECK(cudaDeviceSynchronize());
            if (tiled) cudaFree(tiled);
            WIPA_CUDA_CHECK(cudaMalloc(&tiled, elems * sizeof(h16)));
            tiled_elems = elems;
        }
        WIPA_CUDA_CHECK(cudaMemsetAsync(tiled, 0, elems * sizeof(h16), (cudaStream_t)stream));
        WIPA_TRY(launch_lat_tile(E, 1, tiled, U, T, H, cross_attention_latent_keys(H), (cudaStream_t)stream));
        Ein = tiled;
    }
    if (wide) return launch_cross_attention_latent_wide((const h16*)Qp, Ein, utt_of_seq, (h16*)C, S, H, T, part, part_floats, counters, (cudaStream_t)stream, beams);
    return launch_cross_attention_latent((const h16*)Qp, Ein, layout != 0, U, utt_of_seq, (h16*)C, S, H, T, part, part_floats, counters,
                                         (cudaStream_t)stream, beams);
}

// row-major h16 encoder output [U, T, 64 H] -> the chunk-tiled layout of the latent kernel (layout 2 above); `out` holds
// U * wipa_test_lat_tiled_elems(H, T) h16 elements and must be zeroed by the caller once (padding of the last chunk)
extern "C" long long wipa_test_lat_tiled_elems(int H, int T) {
    return (cross_attention_latent_supported(H) || cross_attention_latent_wide_supported(H)) ? (long long)cross_attention_latent_tiled_elems(H, T) : -1;
}
extern "C" int wipa_test_lat_tile(const void* E, int U, int T, int H, void* out, void* stream) {
    WIPA_CHECK(E && out && (cross_attention_latent_supported(H) || cross_attention_latent_wide_supported(H)), WIPA_EINVAL, "wipa_test_lat_tile: bad argument");
    return launch_lat_tile(E, 1, (h16*)out, U, T, H, cross_attention_latent_keys(H), (cudaStream_t)stream);
}

// absorbed cross-attention queries in one node (gemm_q2.cu): A h16 [S, 64 H], Wq / Wk h16 [64 H, 64 H] row-major
// ([out, in] like nn.Linear), bias f32 [64 H] or NULL -> out h16 [S, H, 64 H] with out[s, h] = Wk_h^T (Wq_h A[s] + bias_h);
// wkt_scratch: h16 [H, 64 H, 64] (holds the per-head transposed Wk)
extern "C" int wipa_test_xlq_fused(const void* A, const void* Wq, const void* Wk, const float* bias, void* wkt_scratch, void* out,
                                   int S, int H, void* stream) {
    WIPA_CHECK(A && Wq && Wk && wkt_scratch && out && S >= 1 && H >= 1 && H % 2 == 0, WIPA_EINVAL, "wipa_test_xlq_fused: bad argument");
    const int d = 64 * H;
    xlat_wkt_kernel<<<dim3(cdiv(d, 4), H), 256, 0, (cudaStream_t)stream>>>((const h16*)Wk, (h16*)wkt_scratch, d);
    WIPA_LAUNCHED();
    return launch_xlq_fused((const h16*)A, (const h16*)Wq, (const h16*)wkt_scratch, bias, nullptr, nullptr, 0, (h16*)out, S, H,
                            (cudaStream_t)stream);
}

// decoder self-attention step alone over a caller-built paged cache: kpool / vpool [page][H][16][64] (h16 or f32),
// block_table int32 [B, bt_stride], *pos_ptr = index of the newest position; q f32 [B, H*64]; out [B, H*64] in the pool type
extern "C" int wipa_test_self_attn(const float* q, const void* kpool, const void* vpool, const int* block_table, int bt_stride,
                                   const int* pos_ptr, void* out, int B, int H, int is_h16, void* stream) {
    WIPA_CHECK(q && kpool && vpool && block_table && pos_ptr && out && B >= 1 && H >= 1, WIPA_EINVAL, "wipa_test_self_attn: bad argument");
    if (is_h16) return launch_self_attention<h16>(q, (const h16*)kpool, (const h16*)vpool, block_table, bt_stride, pos_ptr,
                                                    (h16*)out, B, H, (cudaStream_t)stream);
    return launch_self_attention<float>(q, (const float*)kpool, (const float*)vpool, block_table, bt_stride, pos_ptr, (float*)out,
                                        B, H, (cudaStream_t)stream);
}

// encoder self-attention on h16 device buffers, no conversions (timing): q,k,v h16 [B,H,T,64] -> out h16 [B,T,H*64]
extern "C" int wipa_test_enc_attention_h16(const void* q, const void* k, const void* v, void* out, int B, int H, int T, int tc,
                                            void* stream) {
    if (tc) return launch_enc_attention_tc((const h16*)q, (const h16*)k, (const h16*)v, (h16*)out, B, H, T, (cudaStream_t)stream);
    return launch_enc_attention<h16>((const h16*)q, (const h16*)k, (const h16*)v, (h16*)out, B, H, T, (cudaStream_t)stream);
}

// encoder self-attention alone: q,k,v f32 [B,H,T,64] (q pre-scaled) -> out f32 [B,T,H*64]; use_h16 selects the kernel family
extern "C" int wipa_test_enc_attention(const float* q, const float* k, const float* v, float* out, int B, int H, int T,
                                       int use_h16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!use_h16) return launch_enc_attention<float>(q, k, v, out, B, H, T, st);
    const long long n = (long long)B * H * T * 64;
    h16* buf = nullptr;
    WIPA_CUDA_CHECK(cudaMalloc(&buf, sizeof(h16) * (size_t)n * 4));
    int r = launch_convert(q, buf, n, 1.f, 1, st);
    if (r == WIPA_OK) r = launch_convert(k, buf + n, n, 1.f, 1, st);
    if (r == WIPA_OK) r = launch_convert(v, buf + 2 * n, n, 1.f, 1, st);
    if (r == WIPA_OK) r = use_h16 == 2 ? launch_enc_attention_tc(buf, buf + n, buf + 2 * n, buf + 3 * n, B, H, T, st)
                                        : launch_enc_attention<h16>(buf, buf + n, buf + 2 * n, buf + 3 * n, B, H, T, st);
    if (r == WIPA_OK) r = launch_to_f32(buf + 3 * n, 1, out, n, st);
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return r;
}
