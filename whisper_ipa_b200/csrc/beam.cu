// Beam search bookkeeping on the device, with the semantics of HF's vectorised `_beam_search`
// (HF:generation/utils.py:3076-3400, helpers :2876-3073; early_stopping = False, one EOS id, do_sample = False):
//
//   beam_row_topk_kernel   one CTA per (utterance, beam) row of fp32 logits: log_softmax over the whole vocabulary
//                          (HF applies the suppress processors AFTER log_softmax, :3262-3263), add the running beam
//                          score, keep the row's 2K best continuations            (:3288-3300, _get_top_k_continuations)
//   beam_update_kernel     one CTA per utterance: merge the K rows' candidates into the utterance's top 2K, mark the
//                          ones that hit EOS / max_length, pick the K best unfinished as the next running beams
//                          (:2999-3019), fold newly finished top-K candidates into the K finished slots with the length
//                          penalty (:3021-3073), update the early-stop heuristic (:2876-2921), and — instead of
//                          reordering the KV cache like HF (:3346-3351) — update the per-beam ANCESTRY table: the
//                          self-attention kernel reads position q of beam s from the slot anc[s][q] that wrote it, so
//                          no cached key/value ever moves.
//
// Ties between equal scores are broken towards the lower flat index (torch.topk leaves them unspecified).
#define WIPA_PDL_CLASS 1
#include "common.cuh"

namespace {

constexpr int BEAM_MAX_KEEP = 16;                 // 2 * beams <= 16
constexpr int ROW_THREADS = 512;

struct Cand {
    float v;
    int idx;
};
__device__ __forceinline__ bool better(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

__global__ void __launch_bounds__(ROW_THREADS)
beam_row_topk_kernel(const float* __restrict__ logits, long long ld, int V, const uint32_t* __restrict__ mask_always,
                     const uint32_t* __restrict__ mask_begin, const int* __restrict__ step_ptr,
                     const float* __restrict__ run_score, int keep, float* __restrict__ out_val, int* __restrict__ out_idx) {
    __shared__ float s_red[ROW_THREADS / 32];
    __shared__ double s_redd[ROW_THREADS / 32];
    __shared__ float s_val[ROW_THREADS * BEAM_MAX_KEEP / 4];       // per-warp merged lists (see below)
    __shared__ int s_idx[ROW_THREADS * BEAM_MAX_KEEP / 4];
    __shared__ float s_bcast;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    pdl_launch_dependents();
    const float* x = logits + (size_t)row * ld;
    const bool begin = (step_ptr != nullptr) && (*step_ptr == 0);

    // ---- log_softmax statistics over the full row (fp32 max, fp64 sum) ----------------------------------------------
    float mx = -INFINITY;
    for (int n = tid; n < V; n += ROW_THREADS) mx = fmaxf(mx, x[n]);
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int w = 1; w < ROW_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
        s_bcast = m;
    }
    __syncthreads();
    mx = s_bcast;
    double sum = 0.0;
    for (int n = tid; n < V; n += ROW_THREADS) sum += (double)expf(x[n] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_redd[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < ROW_THREADS / 32; ++w) t += s_redd[w];
        s_bcast = (float)log(t);
    }
    __syncthreads();
    const float logsum = s_bcast;
    const float base = run_score[row];

    // ---- per-thread sorted top-`keep` of lp + running score -------------------------------------------------------------
    Cand best[BEAM_MAX_KEEP];
#pragma unroll
    for (int i = 0; i < BEAM_MAX_KEEP; ++i) { best[i].v = -INFINITY; best[i].idx = 0x7fffffff; }
    for (int n = tid; n < V; n += ROW_THREADS) {
        bool masked = false;
        if (mask_always != nullptr) masked = (mask_always[n >> 5] >> (n & 31)) & 1u;
        if (begin && mask_begin != nullptr) masked = masked || ((mask_begin[n >> 5] >> (n & 31)) & 1u);
        if (masked) continue;                                    // log-prob -inf: never a candidate
        const float v = ((x[n] - mx) - logsum) + base;
        if (better(v, n, best[BEAM_MAX_KEEP - 1].v, best[BEAM_MAX_KEEP - 1].idx)) {
            Cand c{v, n};
#pragma unroll
            for (int i = 0; i < BEAM_MAX_KEEP; ++i) {
                if (better(c.v, c.idx, best[i].v, best[i].idx)) { const Cand t = best[i]; best[i] = c; c = t; }
            }
        }
    }
    // ---- merge: `keep` rounds of (warp argmax over the lanes' list heads), then the same over the warps' lists -------
    // each lane's list is sorted, so a head pointer per lane is enough
    int head = 0;
    for (int r = 0; r < keep; ++r) {
        float v = -INFINITY;
        int idx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < BEAM_MAX_KEEP; ++i) if (i == head) { v = best[i].v; idx = best[i].idx; }
        float bv = v;
        int bi = idx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (bv == v && bi == idx && idx != 0x7fffffff) ++head;   // the unique winner (indices are distinct) advances
        if (lane == 0) { s_val[warp * BEAM_MAX_KEEP + r] = bv; s_idx[warp * BEAM_MAX_KEEP + r] = bi; }
    }
    __syncthreads();
    if (warp == 0) {
        // lane w < n_warps walks warp w's sorted list
        constexpr int NW = ROW_THREADS / 32;
        int h2 = 0;
        for (int r = 0; r < keep; ++r) {
            float v = -INFINITY;
            int idx = 0x7fffffff;
            if (lane < NW && h2 < keep) { v = s_val[lane * BEAM_MAX_KEEP + h2]; idx = s_idx[lane * BEAM_MAX_KEEP + h2]; }
            float bv = v;
            int bi = idx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (bv == v && bi == idx && idx != 0x7fffffff) ++h2;
            if (lane == 0) { out_val[(size_t)row * keep + r] = bv; out_idx[(size_t)row * keep + r] = bi; }
        }
    }
}

// One CTA (128 threads) per utterance.  Sequences are [utt][beam][L]; `flip` selects the current buffer of the
// ping-pong arrays (running sequences, finished sequences, ancestry).
__global__ void __launch_bounds__(128)
beam_update_kernel(BeamState bs, const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int V, int n_utts) {
    __shared__ float c_val[BEAM_MAX_KEEP];
    __shared__ int c_beam[BEAM_MAX_KEEP], c_tok[BEAM_MAX_KEEP], c_hit[BEAM_MAX_KEEP];
    __shared__ int run_src[BEAM_MAX_KEEP], fin_src[BEAM_MAX_KEEP];     // fin_src: < K old finished slot, >= K candidate K + c
    __shared__ float fin_val[BEAM_MAX_KEEP];
    __shared__ int fin_flag[BEAM_MAX_KEEP];
    const int u = blockIdx.x, tid = threadIdx.x;
    const int K = bs.beams, keep = 2 * bs.beams, L = bs.L, P = bs.prompt_len;
    pdl_wait();
    pdl_launch_dependents();
    const int pos = *bs.pos;                                   // position consumed by this step
    const int cur_len = pos + 1;                               // tokens per running sequence so far (HF's cur_len)
    const int flip = *bs.flip;
    const size_t SL = (size_t)n_utts * K * L;
    const int* run_seq = bs.run_seq + flip * SL;
    int* run_seq_n = bs.run_seq + (flip ^ 1) * SL;
    const int* fin_seq = bs.fin_seq + flip * SL;
    int* fin_seq_n = bs.fin_seq + (flip ^ 1) * SL;
    const int* anc = bs.anc + flip * SL;
    int* anc_n = bs.anc + (flip ^ 1) * SL;

    if (tid == 0) {
        // ---- utterance top-2K over the K rows' sorted candidate lists (flat index = beam * V + token) -------------
        int head[BEAM_MAX_KEEP / 2];
        for (int k = 0; k < K; ++k) head[k] = 0;
        for (int r = 0; r < keep; ++r) {
            float bv = -INFINITY;
            long long bflat = 0x7fffffffffffffffLL;
            int bk = 0;
            for (int k = 0; k < K; ++k) {
                if (head[k] >= keep) continue;
                const size_t o = ((size_t)u * K + k) * keep + head[k];
                const float v = cand_val[o];
                const long long flat = (long long)k * V + cand_idx[o];
                if (v > bv || (v == bv && flat < bflat)) { bv = v; bflat = flat; bk = k; }
            }
            c_val[r] = bv;
            c_beam[r] = bk;
            c_tok[r] = cand_idx[((size_t)u * K + bk) * keep + head[bk]];
            ++head[bk];
            c_hit[r] = (c_tok[r] == bs.eot) || (cur_len + 1 >= bs.max_length);
        }
        // ---- next running beams: K best of val + hit * -1e9 (stable) -------------------------------------------------------
        float rv[BEAM_MAX_KEEP];
        bool used[BEAM_MAX_KEEP];
        for (int c = 0; c < keep; ++c) { rv[c] = c_val[c] + (c_hit[c] ? 1.0f : 0.0f) * -1.0e9f; used[c] = false; }
        for (int j = 0; j < K; ++j) {
            int bc = -1;
            for (int c = 0; c < keep; ++c) if (!used[c] && (bc < 0 || rv[c] > rv[bc])) bc = c;
            used[bc] = true;
            run_src[j] = bc;
            bs.run_score_next[(size_t)u * K + j] = rv[bc];
        }
        // ---- finished slots: K best of [old finished scores, penalised candidates] ----------------------------------------
        const float denom = powf((float)(cur_len + 1 - P), bs.length_penalty);
        const bool unsat = bs.unsat[u] != 0;
        float mv[BEAM_MAX_KEEP + BEAM_MAX_KEEP / 2];
        int mf[BEAM_MAX_KEEP + BEAM_MAX_KEEP / 2];
        for (int k = 0; k < K; ++k) { mv[k] = bs.fin_score[(size_t)u * K + k]; mf[k] = bs.fin_done[(size_t)u * K + k]; }
        for (int c = 0; c < keep; ++c) {
            const bool did = c_hit[c] && c < K;
            float s = c_val[c] / denom;
            s += (unsat ? 0.0f : 1.0f) * -1.0e9f;
            s += (did ? 0.0f : 1.0f) * -1.0e9f;
            mv[K + c] = s;
            mf[K + c] = did ? 1 : 0;
        }
        bool taken[BEAM_MAX_KEEP + BEAM_MAX_KEEP / 2];
        for (int i = 0; i < K + keep; ++i) taken[i] = false;
        for (int j = 0; j < K; ++j) {
            int bi = -1;
            for (int i = 0; i < K + keep; ++i) if (!taken[i] && (bi < 0 || mv[i] > mv[bi])) bi = i;
            taken[bi] = true;
            fin_src[j] = bi;
            fin_val[j] = mv[bi];
            fin_flag[j] = mf[bi];
        }
        // ---- early-stop heuristic with the NEW running / finished state (cur_len already advanced) ---------------------------
        const float best_possible = rv[run_src[0]] / powf((float)(cur_len + 1 - P), bs.length_penalty);
        float mn = fin_val[0];
        for (int j = 1; j < K; ++j) mn = fminf(mn, fin_val[j]);
        bool any_better = false;
        for (int j = 0; j < K; ++j) {
            const float worst = fin_flag[j] ? mn : -1.0e9f;
            any_better = any_better || (best_possible > worst);
        }
        const bool unsat_new = unsat && any_better;
        bs.unsat[u] = unsat_new ? 1 : 0;
        if (unsat && !unsat_new) atomicAdd(bs.n_done, 1);         // this utterance can no longer improve
        for (int j = 0; j < K; ++j) {
            bs.fin_score_next[(size_t)u * K + j] = fin_val[j];
            bs.fin_done_next[(size_t)u * K + j] = fin_flag[j];
            bs.cur_tok[(size_t)u * K + j] = c_tok[run_src[j]];
        }
    }
    __syncthreads();
    // ---- move sequences and ancestry rows (all threads) ----------------------------------------------------------------------
    for (int j = 0; j < K; ++j) {
        const int c = run_src[j];
        const int sb = c_beam[c];
        const int* src = run_seq + ((size_t)u * K + sb) * L;
        int* dst = run_seq_n + ((size_t)u * K + j) * L;
        for (int t = tid; t < L; t += blockDim.x) dst[t] = t < cur_len ? src[t] : (t == cur_len ? c_tok[c] : bs.eot);
        const int* asrc = anc + ((size_t)u * K + sb) * L;
        int* adst = anc_n + ((size_t)u * K + j) * L;
        for (int t = tid; t < L; t += blockDim.x) adst[t] = t < pos ? asrc[t] : (u * K + sb);   // position `pos` was written by slot sb
        const int f = fin_src[j];
        int* fdst = fin_seq_n + ((size_t)u * K + j) * L;
        if (f < K) {
            const int* fsrc = fin_seq + ((size_t)u * K + f) * L;
            for (int t = tid; t < L; t += blockDim.x) fdst[t] = fsrc[t];
        } else {
            const int cc = f - K;
            const int* fsrc = run_seq + ((size_t)u * K + c_beam[cc]) * L;
            for (int t = tid; t < L; t += blockDim.x) fdst[t] = t < cur_len ? fsrc[t] : (t == cur_len ? c_tok[cc] : bs.eot);
        }
    }
}

// After every utterance's update: swap the ping-pong buffers, advance the position (single thread).
__global__ void beam_advance_kernel(BeamState bs, int n_seqs) {
    pdl_wait();
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seqs) {
        bs.run_score[i] = bs.run_score_next[i];
        bs.fin_score[i] = bs.fin_score_next[i];
        bs.fin_done[i] = bs.fin_done_next[i];
    }
    if (i == 0) {
        *bs.flip ^= 1;
        const int pos = *bs.pos;
        *bs.pos = pos + 1;
        *bs.step = pos + 1 - (bs.prompt_len - 1);
    }
}

__global__ void beam_init_kernel(BeamState bs, const int* __restrict__ prompt, int n_utts) {
    const int K = bs.beams, L = bs.L;
    const size_t SL = (size_t)n_utts * K * L;
    const size_t n = SL;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % L);
        const int s = (int)(i / L);
        const int tok = t < bs.prompt_len ? prompt[t] : bs.eot;
        bs.run_seq[i] = tok; bs.run_seq[SL + i] = tok;
        bs.fin_seq[i] = tok; bs.fin_seq[SL + i] = tok;
        bs.anc[i] = s; bs.anc[SL + i] = s;
        if (t == 0) {
            bs.run_score[s] = (s % K == 0) ? 0.0f : -1.0e9f;
            bs.fin_score[s] = -1.0e9f;
            bs.fin_done[s] = 0;
        }
        if (i < (size_t)n_utts) bs.unsat[i] = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *bs.flip = 0;
}

// best finished hypothesis of every utterance -> out_ids [n_utts, max_new] (EOT padded), out_len [n_utts]
__global__ void beam_finish_kernel(BeamState bs, int n_utts, int max_new, int* __restrict__ out_ids, int* __restrict__ out_len) {
    const int u = blockIdx.x;
    const int K = bs.beams, L = bs.L;
    const size_t SL = (size_t)n_utts * K * L;
    const int* seq = bs.fin_seq + (size_t)(*bs.flip) * SL + (size_t)u * K * L;      // slot 0 = best (scores sorted descending)
    __shared__ int s_len;
    if (threadIdx.x == 0) {
        int n = 0;
        while (n < max_new && seq[bs.prompt_len + n] != bs.eot) ++n;
        s_len = n;
        out_len[u] = n;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < max_new; t += blockDim.x) out_ids[(size_t)u * max_new + t] = t < s_len ? seq[bs.prompt_len + t] : bs.eot;
}

}  // namespace

int launch_beam_row_topk(const float* logits, long long ld, int V, const uint32_t* mask_always, const uint32_t* mask_begin,
                         const int* step_ptr, const float* run_score, int keep, float* out_val, int* out_idx, int n_rows,
                         cudaStream_t st) {
    WIPA_CHECK(keep >= 2 && keep <= BEAM_MAX_KEEP, WIPA_EINVAL, "beam search: 2 * beams must be <= %d", BEAM_MAX_KEEP);
    WIPA_CUDA_CHECK(wipa_launch(beam_row_topk_kernel, dim3(n_rows), dim3(ROW_THREADS), (size_t)0, st, logits, ld, V, mask_always,
                                mask_begin, step_ptr, run_score, keep, out_val, out_idx));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

int launch_beam_update(const BeamState& bs, const float* cand_val, const int* cand_idx, int V, int n_utts, cudaStream_t st) {
    WIPA_CUDA_CHECK(wipa_launch(beam_update_kernel, dim3(n_utts), dim3(128), (size_t)0, st, bs, cand_val, cand_idx, V, n_utts));
    WIPA_LAUNCHED();
    const int n_seqs = n_utts * bs.beams;
    WIPA_CUDA_CHECK(wipa_launch(beam_advance_kernel, dim3(cdiv(n_seqs, 256)), dim3(256), (size_t)0, st, bs, n_seqs));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

int launch_beam_init(const BeamState& bs, const int* prompt_dev, int n_utts, cudaStream_t st) {
    beam_init_kernel<<<148, 256, 0, st>>>(bs, prompt_dev, n_utts);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

int launch_beam_finish(const BeamState& bs, int n_utts, int max_new, int* out_ids, int* out_len, cudaStream_t st) {
    beam_finish_kernel<<<n_utts, 128, 0, st>>>(bs, n_utts, max_new, out_ids, out_len);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
