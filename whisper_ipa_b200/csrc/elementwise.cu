// Small bandwidth-bound kernels around the GEMMs: LayerNorm, token/position embedding, mel layout change,
// weight conversion, fp32-path vocabulary argmax and the on-device greedy bookkeeping (token feedback, EOT tracking).
#include "common.cuh"

// ================================================================================================
// LayerNorm (eps 1e-5, HF:models/whisper/modeling_whisper.py:369,430-432): one warp per row, the row
// lives in registers (d = 128 * NV, NV float4 per lane), two-pass mean / variance in fp32.
// ================================================================================================
template <typename T, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 T* __restrict__ out, int M) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap
    if (row >= M) return;
    constexpr int d = NV * 128;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[i * 32 + lane];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, dd = v[i].w - mean;
        q += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        const float4 wv = *reinterpret_cast<const float4*>(w + c0);
        const float4 bv = *reinterpret_cast<const float4*>(b + c0);
        float r[4] = {(v[i].x - mean) * rstd * wv.x + bv.x, (v[i].y - mean) * rstd * wv.y + bv.y,
                      (v[i].z - mean) * rstd * wv.z + bv.z, (v[i].w - mean) * rstd * wv.w + bv.w};
        store_group<4>(out, sizeof(T) == 2, (long long)row * d + c0, r, true);
    }
}

// The encoder's final LayerNorm when the decoder attends over the encoder output itself: same arithmetic, h16 result written in
// the chunk-tiled, pre-swizzled layout the latent cross-attention kernel streams (common.cuh lat_tile_offset).
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_lat_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, h16* __restrict__ out,
                     int M, int T, int u0, int keys, size_t utt_elems) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    constexpr int d = NV * 128;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[i * 32 + lane];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, dd = v[i].w - mean;
        q += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + 1e-5f);
    const int u = row / T, t = row - u * T;
    h16* base = out + (size_t)(u0 + u) * utt_elems;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        const float4 wv = *reinterpret_cast<const float4*>(w + c0);
        const float4 bv = *reinterpret_cast<const float4*>(b + c0);
        uint2 pk;
        pk.x = pack_h16x2((v[i].x - mean) * rstd * wv.x + bv.x, (v[i].y - mean) * rstd * wv.y + bv.y);
        pk.y = pack_h16x2((v[i].z - mean) * rstd * wv.z + bv.z, (v[i].w - mean) * rstd * wv.w + bv.w);
        *reinterpret_cast<uint2*>(base + lat_tile_offset(t, c0, d / 64, keys)) = pk;
    }
}

int launch_layernorm_lat(const float* x, const float* w, const float* b, h16* out_tiled, int M, int d, int T, int u0, int keys,
                         cudaStream_t st) {
    WIPA_CHECK(d % 128 == 0 && d / 128 <= 10, WIPA_EUNSUPPORTED, "layernorm: d=%d must be 128*k, k<=10", d);
    if (M == 0) return WIPA_OK;
    const int grid = cdiv(M, 8);
    const size_t utt = (size_t)cdiv(T, keys) * keys * d;
    switch (d / 128) {
#define LN_CASE(NV) case NV: layernorm_lat_kernel<NV><<<grid, 256, 0, st>>>(x, w, b, out_tiled, M, T, u0, keys, utt); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8) LN_CASE(9) LN_CASE(10)
#undef LN_CASE
    }
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// row-major encoder output [U, T, d] (f32 or h16) -> chunk-tiled h16 (audio features handed in from outside; kernel tests)
__global__ void __launch_bounds__(256)
lat_tile_kernel(const void* __restrict__ src, int src_is_h16, h16* __restrict__ out, int T, int d, int keys, size_t utt_elems) {
    const int u = blockIdx.y, t = blockIdx.x;
    h16* base = out + (size_t)u * utt_elems;
    const size_t in0 = ((size_t)u * T + t) * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float v = src_is_h16 ? h16_to_f32(reinterpret_cast<const h16*>(src)[in0 + c]) : reinterpret_cast<const float*>(src)[in0 + c];
        base[lat_tile_offset(t, c, d / 64, keys)] = f32_to_h16(v);
    }
}

int launch_lat_tile(const void* src, int src_is_h16, h16* out_tiled, int U, int T, int H, int keys, cudaStream_t st) {
    if (U == 0) return WIPA_OK;
    const int d = 64 * H;
    lat_tile_kernel<<<dim3(T, U), 256, 0, st>>>(src, src_is_h16, out_tiled, T, d, keys, (size_t)cdiv(T, keys) * keys * d);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

template <typename T>
int launch_layernorm(const float* x, const float* w, const float* b, T* out, int M, int d, cudaStream_t st) {
    WIPA_CHECK(d % 128 == 0 && d / 128 <= 10, WIPA_EUNSUPPORTED, "layernorm: d=%d must be 128*k, k<=10", d);
    if (M == 0) return WIPA_OK;
    const int grid = cdiv(M, 8);
    switch (d / 128) {
#define LN_CASE(NV) case NV: WIPA_CUDA_CHECK(wipa_launch(layernorm_kernel<T, NV>, dim3(grid), dim3(256), (size_t)0, st, x, w, b, out, M)); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8) LN_CASE(9) LN_CASE(10)
#undef LN_CASE
    }
    WIPA_LAUNCHED();
    return WIPA_OK;
}
template int launch_layernorm<float>(const float*, const float*, const float*, float*, int, int, cudaStream_t);
template int launch_layernorm<h16>(const float*, const float*, const float*, h16*, int, int, cudaStream_t);

// ================================================================================================
// mel [B, C, 3000] f32 (HF layout) -> rows [B, 3002, C] in T, written at rows 1..3000; rows 0 and 3001 are the
// conv padding and are zeroed here too, so conv1d(k=3, p=1) becomes a GEMM over overlapping rows.
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
mel_to_rows_kernel(const float* __restrict__ mel, T* __restrict__ rows, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, t = t0 + tx;
        tile[j][tx] = (c < C && t < WIPA_N_FRAMES) ? mel[((size_t)b * C + c) * WIPA_N_FRAMES + t] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, c = c0 + tx;
        if (t < WIPA_N_FRAMES && c < C) rows[((size_t)b * (WIPA_N_FRAMES + 2) + 1 + t) * C + c] = from_f32<T>(tile[tx][j]);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int c = threadIdx.x; c < C; c += 256) {
            rows[((size_t)b * (WIPA_N_FRAMES + 2)) * C + c] = from_f32<T>(0.f);
            rows[((size_t)b * (WIPA_N_FRAMES + 2) + WIPA_N_FRAMES + 1) * C + c] = from_f32<T>(0.f);
        }
    }
}

template <typename T>
int launch_mel_to_rows(const float* mel, T* rows, int B, int C, cudaStream_t st) {
    if (B == 0) return WIPA_OK;
    dim3 grid(cdiv(WIPA_N_FRAMES, 32), cdiv(C, 32), B);
    mel_to_rows_kernel<T><<<grid, 256, 0, st>>>(mel, rows, C);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
template int launch_mel_to_rows<float>(const float*, float*, int, int, cudaStream_t);
template int launch_mel_to_rows<h16>(const float*, h16*, int, int, cudaStream_t);

// ================================================================================================
// decoder input: x[b] = embed_tokens[tok[b]] + embed_positions[*pos]   (HF:...modeling_whisper.py:737-760)
// ================================================================================================
template <typename T>
__global__ void embed_kernel(const T* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                             const int* __restrict__ tok, const int* __restrict__ pos_ptr, float* __restrict__ x, int d) {
    const int b = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap
    const int p = *pos_ptr;
    const T* te = tok_emb + (size_t)tok[b] * d;
    const float* pe = pos_emb + (size_t)p * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[(size_t)b * d + i] = to_f32(te[i]) + pe[i];
}

template <typename T>
int launch_embed(const T* tok_emb, const float* pos_emb, const int* tok, const int* pos_ptr, float* x, int Bs, int d,
                 cudaStream_t st) {
    WIPA_CUDA_CHECK(wipa_launch(embed_kernel<T>, dim3(Bs), dim3(256), (size_t)0, st, tok_emb, pos_emb, tok, pos_ptr, x, d));
    WIPA_LAUNCHED();
    return WIPA_OK;
}
// The same for the folded-LayerNorm decode path (common.cuh): besides x it leaves x rounded to h16 (the first GEMM's A
// operand) and the (mean, M2) of every 32-column piece of the row.  Warp w owns pieces w, w + 8, ...; lane = column in the piece.
__global__ void __launch_bounds__(256)
embed_lnf_kernel(const h16* __restrict__ tok_emb, const float* __restrict__ pos_emb, const int* __restrict__ tok,
                 const int* __restrict__ pos_ptr, float* __restrict__ x, h16* __restrict__ x16, float* __restrict__ stats, int d) {
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    const int p = *pos_ptr;
    const h16* te = tok_emb + (size_t)tok[b] * d;
    const float* pe = pos_emb + (size_t)p * d;
    const int nt = d / WIPA_LN_PIECE;
    for (int t = warp; t < nt; t += 8) {
        const int i = t * WIPA_LN_PIECE + lane;
        const float v = h16_to_f32(te[i]) + pe[i];
        x[(size_t)b * d + i] = v;
        x16[(size_t)b * d + i] = f32_to_h16(v);
        const float mean = warp_sum(v) * (1.0f / WIPA_LN_PIECE);
        const float dlt = v - mean;
        const float m2 = warp_sum(dlt * dlt);
        if (lane == 0) *reinterpret_cast<float2*>(stats + ((size_t)b * nt + t) * 2) = make_float2(mean, m2);
    }
}

int launch_embed_lnf(const h16* tok_emb, const float* pos_emb, const int* tok, const int* pos_ptr, float* x, h16* x16, float* stats,
                     int Bs, int d, cudaStream_t st) {
    WIPA_CUDA_CHECK(wipa_launch(embed_lnf_kernel, dim3(Bs), dim3(256), (size_t)0, st, tok_emb, pos_emb, tok, pos_ptr, x, x16, stats, d));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

template int launch_embed<float>(const float*, const float*, const int*, const int*, float*, int, int, cudaStream_t);
template int launch_embed<h16>(const h16*, const float*, const int*, const int*, float*, int, int, cudaStream_t);

// ================================================================================================
// weight import: fp32 state_dict tensor -> context storage (fp32 or h16), optional exact scale (q * 2^-3)
// ================================================================================================
__global__ void convert_kernel(const float* __restrict__ src, void* __restrict__ dst, long long n, float scale, int to_h16) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = src[i] * scale;
        if (to_h16) reinterpret_cast<h16*>(dst)[i] = f32_to_h16(v);
        else reinterpret_cast<float*>(dst)[i] = v;
    }
}

int launch_convert(const float* src, void* dst, long long n, float scale, int to_h16, cudaStream_t st) {
    if (n == 0) return WIPA_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    convert_kernel<<<(int)blocks, 256, 0, st>>>(src, dst, n, scale, to_h16);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// conv1d weight [N, C, 3] -> GEMM weight [N, 3*C] with k = tap * C + c (matches rows of [x[t-1] | x[t] | x[t+1]])
__global__ void conv_weight_kernel(const float* __restrict__ src, void* __restrict__ dst, int N, int C, int to_h16) {
    const long long total = (long long)N * C * 3;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int n = (int)(i / (3 * C));
        const int r = (int)(i - (long long)n * 3 * C);
        const int tap = r / C, c = r - tap * C;
        const float v = src[((long long)n * C + c) * 3 + tap];
        if (to_h16) reinterpret_cast<h16*>(dst)[i] = f32_to_h16(v);
        else reinterpret_cast<float*>(dst)[i] = v;
    }
}

int launch_conv_weight(const float* src, void* dst, int N, int C, int to_h16, cudaStream_t st) {
    long long blocks = ((long long)N * C * 3 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    conv_weight_kernel<<<(int)blocks, 256, 0, st>>>(src, dst, N, C, to_h16);
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// ================================================================================================
// fp32 path: argmax over a materialised logits row with the suppress masks (first index wins ties, like torch.argmax)
// ================================================================================================
__global__ void __launch_bounds__(256)
row_argmax_kernel(const float* __restrict__ logits, int V, const uint32_t* __restrict__ mask_always,
                  const uint32_t* __restrict__ mask_begin, const int* __restrict__ step_ptr, float* __restrict__ pmax,
                  int* __restrict__ pidx) {
    __shared__ float smax[8];
    __shared__ int sidx[8];
    const int b = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap
    const float* row = logits + (size_t)b * V;
    const bool begin = (step_ptr != nullptr) && (*step_ptr == 0);
    float best = -INFINITY;
    int best_n = 0x7fffffff;
    for (int n = threadIdx.x; n < V; n += 256) {
        bool m = false;
        if (mask_always) m = (mask_always[n >> 5] >> (n & 31)) & 1u;
        if (begin && mask_begin) m = m || ((mask_begin[n >> 5] >> (n & 31)) & 1u);
        const float v = row[n];
        if (!m && v > best) { best = v; best_n = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
        if (ov > best || (ov == best && on < best_n)) { best = ov; best_n = on; }
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = best; sidx[threadIdx.x >> 5] = best_n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (smax[w] > best || (smax[w] == best && sidx[w] < best_n)) { best = smax[w]; best_n = sidx[w]; }
        pmax[b] = best;
        pidx[b] = best_n;
    }
}

int launch_row_argmax(const float* logits, int Bs, int V, const uint32_t* mask_always, const uint32_t* mask_begin,
                      const int* step_ptr, float* pmax, int* pidx, cudaStream_t st) {
    WIPA_CUDA_CHECK(wipa_launch(row_argmax_kernel, dim3(Bs), dim3(256), (size_t)0, st, logits, V, mask_always, mask_begin, step_ptr, pmax, pidx));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// ================================================================================================
// greedy bookkeeping, one warp per sequence: reduce the per-tile (max, argmax) partials, apply HF's finished-row
// rule (HF:generation/utils.py:2796-2797: finished rows emit pad = EOT), record the token, feed it back, and
// advance the shared position.  While *pos + 1 < n_forced the next token is the teacher-forced prompt token.
// ================================================================================================
// One warp per sequence, 8 sequences per CTA.  Every CTA reads the step's position first; the LAST CTA to finish (ticket in
// *ds.ticket, self-resetting) advances it, so no CTA can see the incremented value.  (A single 1024-thread CTA walking all
// rows cost 16 us per decode step at 256 sequences x 406 vocabulary pieces.)
__global__ void __launch_bounds__(256)
greedy_finalize_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx, int n_tiles, DecodeState ds, int Bs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap
    const int pos = *ds.pos;
    const int i = pos - (ds.n_forced - 1);            // index of the token sampled at this step
    const int b = blockIdx.x * 8 + warp;
    if (b < Bs) {
        if (i < 0) {
            if (lane == 0) ds.cur_tok[b] = ds.forced[(size_t)b * ds.n_forced + pos + 1];
        } else {
            float best = -INFINITY;
            int best_n = 0x7fffffff;
            for (int t = lane; t < n_tiles; t += 32) {
                const float v = pmax[(size_t)b * n_tiles + t];
                const int n = pidx[(size_t)b * n_tiles + t];
                if (v > best || (v == best && n < best_n)) { best = v; best_n = n; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
                if (ov > best || (ov == best && on < best_n)) { best = ov; best_n = on; }
            }
            if (lane == 0) {
                int tok = best_n;
                if (ds.done[b]) tok = ds.eot;
                else if (tok == ds.eot) {
                    ds.done[b] = 1;
                    ds.out_len[b] = i;
                    atomicAdd(ds.n_done, 1);
                }
                if (i < ds.max_new) ds.out_ids[(size_t)b * ds.max_new + i] = tok;
                ds.cur_tok[b] = tok;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(ds.ticket, 1);
        if (t == (int)gridDim.x - 1) {               // every CTA has read `pos` and written its rows
            *ds.ticket = 0;
            *ds.pos = pos + 1;
            *ds.step = i + 1;
        }
    }
}

int launch_greedy_finalize(const float* pmax, const int* pidx, int n_tiles, DecodeState ds, int Bs, cudaStream_t st) {
    WIPA_CUDA_CHECK(wipa_launch(greedy_finalize_kernel, dim3(cdiv(Bs, 8)), dim3(256), (size_t)0, st, pmax, pidx, n_tiles, ds, Bs));
    WIPA_LAUNCHED();
    return WIPA_OK;
}
