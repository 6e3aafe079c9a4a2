// Attention kernels.
//   enc_attention_kernel   : encoder self-attention (non-causal, 1500 x 1500 x 64 per head), flash-style
//                            online softmax, fp32 math on CUDA cores.  This is the fp32-path kernel
//                            (HF:models/whisper/modeling_whisper.py:284-357 with scaling folded into q).
//                            The h16 path uses the tcgen05 flash kernel in attn_tc.cu.
//   self_attention_kernel  : one decode step of decoder self-attention over the paged KV cache (one CTA of four warps
//                            per (seq, head); optional beam-ancestry indirection).
//   cross_attention_stream_kernel : one decode step of cross-attention against the cached encoder K/V — the dominant HBM
//                            stream of the whole decode: persistent stream-K streamer (default).
//   cross_attention_kernel : the earlier one-CTA-per-chunk version, kept for A/B runs (WIPA_CA_LEGACY=1).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

// ================================================================================================
// encoder self-attention, SIMT
// ================================================================================================
#define EA_BQ 64
#define EA_BK 64
#define EA_LD (64 + 4)

template <typename T>
__device__ __forceinline__ void load16_as_float(const T* p, float* out);   // 16 consecutive elements
template <>
__device__ __forceinline__ void load16_as_float<float>(const float* p, float* out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
        out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
}
template <>
__device__ __forceinline__ void load16_as_float<h16>(const h16* p, float* out) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint4 u = *reinterpret_cast<const uint4*>(p + 8 * i);
        const h16x2* h = reinterpret_cast<const h16x2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = h16x2_to_f2(h[j]);
            out[8 * i + 2 * j] = f.x; out[8 * i + 2 * j + 1] = f.y;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
enc_attention_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ out,
                     int H, int Tq) {
    extern __shared__ __align__(16) float ea_smem[];
    float* Qt = ea_smem;                        // [64 e][EA_LD q]
    float* Kt = Qt + 64 * EA_LD;                // [64 e][EA_LD key]
    float* Vs = Kt + 64 * EA_LD;                // [64 key][EA_LD e]
    float* Pt = Vs + 64 * EA_LD;                // [64 key][EA_LD q]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int bh = blockIdx.y;
    const int q0 = blockIdx.x * EA_BQ;
    const T* qb = q + (size_t)bh * Tq * 64;
    const T* kb = k + (size_t)bh * Tq * 64;
    const T* vb = v + (size_t)bh * Tq * 64;

    const int lr = tid >> 2;                    // row within a 64-row tile
    const int le = (tid & 3) * 16;              // 16-element chunk
    {
        float tmp[16];
        const int t = q0 + lr;
        if (t < Tq) load16_as_float<T>(qb + (size_t)t * 64 + le, tmp);
        else {
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) Qt[(le + i) * EA_LD + lr] = tmp[i];
    }

    float m_run[4], l_run[4], o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < Tq; k0 += EA_BK) {
        __syncthreads();                         // previous tile fully consumed (also covers the Q store)
        {
            float tk[16], tv[16];
            const int t = k0 + lr;
            if (t < Tq) {
                load16_as_float<T>(kb + (size_t)t * 64 + le, tk);
                load16_as_float<T>(vb + (size_t)t * 64 + le, tv);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { tk[i] = 0.f; tv[i] = 0.f; }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) Kt[(le + i) * EA_LD + lr] = tk[i];
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(&Vs[lr * EA_LD + le + i]) = make_float4(tv[i], tv[i + 1], tv[i + 2], tv[i + 3]);
        }
        __syncthreads();

        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int e = 0; e < 64; ++e) {
            const float4 qv = *reinterpret_cast<const float4*>(&Qt[e * EA_LD + ty * 4]);
            const float4 kv = *reinterpret_cast<const float4*>(&Kt[e * EA_LD + tx * 4]);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
            const float ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
        }
        // mask the key tail, online softmax over the 16 lanes that share a query row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx * 4 + j >= Tq) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);
            const float alpha = expf(m_run[i] - m_new);          // exp(-inf) = 0 on the first tile
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - m_new);
                s[i][j] = p;
                rs += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_run[i] = l_run[i] * alpha + rs;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(&Pt[(tx * 4 + j) * EA_LD + ty * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < EA_BK; ++kk) {
            const float4 pv = *reinterpret_cast<const float4*>(&Pt[kk * EA_LD + ty * 4]);
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[kk * EA_LD + tx * 4]);
            const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
            const float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pa[i], va[j], o[i][j]);
        }
    }
    const int b = bh / H, h = bh - b * H;
    const int d = H * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = q0 + ty * 4 + i;
        if (t < Tq) {
            const float inv = 1.0f / l_run[i];
            float r[4] = {o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv};
            store_group<4>(out, sizeof(T) == 2, ((size_t)b * Tq + t) * d + h * 64 + tx * 4, r, true);
        }
    }
}

template <typename T>
int launch_enc_attention(const T* q, const T* k, const T* v, T* out, int B, int H, int Tq, cudaStream_t st) {
    const size_t smem = sizeof(float) * 4 * 64 * EA_LD;
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(enc_attention_kernel<T>, smem, attr));
    dim3 grid(cdiv(Tq, EA_BQ), B * H);
    enc_attention_kernel<T><<<grid, 256, smem, st>>>(q, k, v, out, H, Tq);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
template int launch_enc_attention<float>(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
template int launch_enc_attention<h16>(const h16*, const h16*, const h16*, h16*, int, int, int, cudaStream_t);

// ================================================================================================
// decoder self-attention, one new token per sequence, paged KV
// ================================================================================================
template <typename T>
__device__ __forceinline__ float dot64(const T* kp, const float* qs);
template <>
__device__ __forceinline__ float dot64<float>(const float* kp, const float* qs) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kp + i);
        acc = fmaf(kv.x, qs[i], acc); acc = fmaf(kv.y, qs[i + 1], acc);
        acc = fmaf(kv.z, qs[i + 2], acc); acc = fmaf(kv.w, qs[i + 3], acc);
    }
    return acc;
}
template <>
__device__ __forceinline__ float dot64<h16>(const h16* kp, const float* qs) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(kp + i);
        const h16x2* h = reinterpret_cast<const h16x2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = h16x2_to_f2(h[j]);
            acc = fmaf(f.x, qs[i + 2 * j], acc);
            acc = fmaf(f.y, qs[i + 2 * j + 1], acc);
        }
    }
    return acc;
}

// ---- helpers shared by the decode-step attention kernels -------------------------------------------------
#define CA_THREADS 256
#define CA_WARPS (CA_THREADS / 32)
#define CA_PART 66                     // floats per partial: m, l, o[64]

template <typename T> struct CaCfg;
template <> struct CaCfg<float> { static constexpr int VEC = 4, LPK = 16; };   // lanes per key row (16 B each)
template <> struct CaCfg<h16> { static constexpr int VEC = 8, LPK = 8; };

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float* f);
template <>
__device__ __forceinline__ void unpack16<float>(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
}
template <>
__device__ __forceinline__ void unpack16<h16>(const uint4& u, float* f) {
    const h16x2* h = reinterpret_cast<const h16x2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = h16x2_to_f2(h[j]);
        f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
}

__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const h16* p) { return h16x2_to_f2(*reinterpret_cast<const h16x2*>(p)); }

// One CTA of SA_WARPS (4) warps per (sequence, head); warp w takes pages w, w+4, ... of the sequence's block table, so a
// sequence of 112 positions is 2 pages (one or two memory round trips) per warp instead of 7 for one warp (8 warps per
// CTA measured slower: 264 vs 226 us per decode step at B=256):
//   per page  all loads are issued up front: the K rows as 16-byte pieces (LPK lanes per key, KPW keys per
//             instruction, every request a full 128-byte line) and the 16 V rows (lane l owns output dims 2l, 2l+1);
//             scores by an LPK-lane shuffle reduction, warp-local online softmax (m, l, o) across the warp's pages
//   merge     the warp partials are combined through shared memory
// (an earlier fully unrolled one-warp version was 125 KB of SASS and thrashed the instruction cache)
template <typename T> struct VRaw;
template <> struct VRaw<float> { typedef float2 type; };
template <> struct VRaw<h16> { typedef h16x2 type; };
__device__ __forceinline__ float2 v_to_f2(float2 v) { return v; }
__device__ __forceinline__ float2 v_to_f2(h16x2 v) { return h16x2_to_f2(v); }
__device__ __forceinline__ void st_v2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st_v2(h16* p, float a, float b) { *reinterpret_cast<h16x2*>(p) = f2_to_h16x2(a, b); }

#define SA_WARPS 4
template <typename T, bool ANC>
__global__ void __launch_bounds__(SA_WARPS * 32, sizeof(T) == 2 ? 8 : 4)
self_attention_kernel(const float* __restrict__ q, const T* __restrict__ kpool, const T* __restrict__ vpool,
                      const int* __restrict__ block_table, int bt_stride, const int* __restrict__ pos_ptr,
                      T* __restrict__ out, int H, const int* __restrict__ anc_base, const int* __restrict__ flip_ptr, int anc_L) {
    using C = CaCfg<T>;
    constexpr int KPW = 32 / C::LPK;                               // keys per warp instruction
    constexpr int ITERS = WIPA_PAGE / KPW;
    typedef typename VRaw<T>::type vraw;
    __shared__ float part[SA_WARPS][CA_PART];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = blockIdx.x;
    const int b = pair / H, h = pair - b * H;
    const int d = H * 64;
    const int grp = lane / C::LPK, li = lane % C::LPK;
    const int* bt = block_table + (size_t)b * bt_stride;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap
    // the warp's first page id does not depend on the position: fetch it next to the position instead of behind it - one
    // dependent round trip less in a kernel whose floor is waves x latency chain (measured at 512 sequences x 12 heads: 15 us
    // at 8 positions whatever the CTA shape - 2 or 4 heads per CTA, 10 - 16 CTAs per SM with spills were all slower; 36.8 ->
    // 36.0 us at 114 positions with this).  Entries past the sequence's pages are valid ints and are not used.
    const int page_first = warp < bt_stride ? bt[warp] : 0;
    const int len = *pos_ptr + 1;
    const int npages = (len + WIPA_PAGE - 1) / WIPA_PAGE;
    // beam search: position p of this sequence lives in the pages of slot anc[p] (p < len - 1); the newest position is
    // always the sequence's own.  Greedy decoding passes anc_base = nullptr (every position is the sequence's own).
    const int* anc = ANC ? anc_base + ((size_t)(*flip_ptr) * gridDim.x / H + b) * anc_L : nullptr;
    float qv[C::VEC];
    {
        const float4* qp = reinterpret_cast<const float4*>(q + (size_t)b * d + h * 64 + li * C::VEC);
#pragma unroll
        for (int i = 0; i < C::VEC / 4; ++i) {
            const float4 t = qp[i];
            qv[4 * i] = t.x; qv[4 * i + 1] = t.y; qv[4 * i + 2] = t.z; qv[4 * i + 3] = t.w;
        }
    }
    float m_run = -INFINITY, l_run = 0.f, a0 = 0.f, a1 = 0.f;
    for (int pg = warp; pg < npages; pg += SA_WARPS) {
        const int page = pg == warp ? page_first : bt[pg];
        const int nkeys = min(WIPA_PAGE, len - pg * WIPA_PAGE);
        const size_t base = ((size_t)page * H + h) * WIPA_PAGE * 64;
        auto key_base = [&](int key) -> size_t {                   // element offset of row `key` of this page
            if (!ANC) return base + (size_t)key * 64;
            const int p = pg * WIPA_PAGE + key;
            if (p == len - 1) return base + (size_t)key * 64;
            const int pg2 = block_table[(size_t)anc[p] * bt_stride + pg];
            return (((size_t)pg2 * H + h) * WIPA_PAGE + key) * 64;
        };
        uint4 kraw[ITERS];
        vraw vr[WIPA_PAGE];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int key = it * KPW + grp;
            kraw[it] = make_uint4(0u, 0u, 0u, 0u);
            if (key < nkeys) kraw[it] = *reinterpret_cast<const uint4*>(kpool + key_base(key) + li * C::VEC);
        }
#pragma unroll
        for (int j = 0; j < WIPA_PAGE; ++j) {
            if (j < nkeys) vr[j] = *reinterpret_cast<const vraw*>(vpool + key_base(j) + 2 * lane);
        }
        float sc[ITERS];
        float mw = -INFINITY;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            float f[C::VEC];
            unpack16<T>(kraw[it], f);
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < C::VEC; ++i) v = fmaf(f[i], qv[i], v);
#pragma unroll
            for (int off = C::LPK / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (it * KPW + grp >= nkeys) v = -INFINITY;
            sc[it] = v;
            mw = fmaxf(mw, v);
        }
#pragma unroll
        for (int off = C::LPK; off < 32; off <<= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, off));
        const float m_new = fmaxf(m_run, mw);                      // finite: the page holds >= 1 key
        const float alpha = expf(m_run - m_new);                   // 0 on the warp's first page
        float lsum = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            sc[it] = expf(sc[it] - m_new);
            lsum += sc[it];
        }
#pragma unroll
        for (int off = C::LPK; off < 32; off <<= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        l_run = fmaf(l_run, alpha, lsum);
        a0 *= alpha; a1 *= alpha;
        m_run = m_new;
#pragma unroll
        for (int j = 0; j < WIPA_PAGE; ++j) {
            const float p = __shfl_sync(0xffffffffu, sc[j / KPW], (j % KPW) * C::LPK);
            if (j < nkeys) {
                const float2 v2 = v_to_f2(vr[j]);
                a0 = fmaf(p, v2.x, a0);
                a1 = fmaf(p, v2.y, a1);
            }
        }
    }
    part[warp][2 + 2 * lane] = a0;
    part[warp][3 + 2 * lane] = a1;
    if (lane == 0) { part[warp][0] = m_run; part[warp][1] = l_run; }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int e = threadIdx.x;
        float M = part[0][0];                                      // warp 0 always has page 0: finite
#pragma unroll
        for (int w = 1; w < SA_WARPS; ++w) M = fmaxf(M, part[w][0]);
        float L = 0.f, o = 0.f;
#pragma unroll
        for (int w = 0; w < SA_WARPS; ++w) {
            const float wgt = expf(part[w][0] - M);
            L = fmaf(part[w][1], wgt, L);
            o = fmaf(part[w][2 + e], wgt, o);
        }
        out[(size_t)b * d + h * 64 + e] = from_f32<T>(o / L);
    }
}

template <typename T>
int launch_self_attention(const float* q, const T* kpool, const T* vpool, const int* block_table, int bt_stride,
                          const int* pos_ptr, T* out, int Bs, int H, cudaStream_t st, const int* anc_base, const int* flip_ptr,
                          int anc_L) {
    if (anc_base != nullptr)
        WIPA_CUDA_CHECK(wipa_launch_c(2, self_attention_kernel<T, true>, dim3(Bs * H), dim3(SA_WARPS * 32), (size_t)0, st, q, kpool, vpool,
                                      block_table, bt_stride, pos_ptr, out, H, anc_base, flip_ptr, anc_L));
    else
        WIPA_CUDA_CHECK(wipa_launch_c(2, self_attention_kernel<T, false>, dim3(Bs * H), dim3(SA_WARPS * 32), (size_t)0, st, q, kpool, vpool,
                                      block_table, bt_stride, pos_ptr, out, H, anc_base, flip_ptr, anc_L));
    WIPA_LAUNCHED();
    return WIPA_OK;
}
template int launch_self_attention<float>(const float*, const float*, const float*, const int*, int, const int*, float*, int, int, cudaStream_t, const int*, const int*, int);
template int launch_self_attention<h16>(const float*, const h16*, const h16*, const int*, int, const int*, h16*, int, int, cudaStream_t, const int*, const int*, int);

// ================================================================================================
// decoder cross-attention, one query per (sequence, head) against 1500 cached encoder keys/values
// ================================================================================================
// This kernel moves ~92 % of the decode step's bytes at batch 64, so it is built as a pure
// HBM streamer.  K/V of one (utterance, head) are contiguous [1500][64].  The 1500 keys are split into n_split
// chunks; CTA (chunk, head, seq) issues two cp.async.bulk copies (its K chunk and its V chunk, each one contiguous
// block) into shared memory behind two mbarriers and only then computes: scores + local softmax from the K
// buffer while V is still landing, then P.V.  Several CTAs are resident per SM (24-48 KB of loads in flight
// each), which is what keeps HBM busy; nothing is read twice.  Each CTA emits a flash-decoding partial
// (max, sum, o[64]); the last CTA to finish a (seq, head) — found with one atomic ticket — merges the partials in
// chunk order (deterministic) and writes the head's output.
template <typename T>
__global__ void __launch_bounds__(CA_THREADS)
cross_attention_kernel(const float* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                       const int* __restrict__ utt_of_seq, T* __restrict__ out, float* __restrict__ part,
                       int* __restrict__ counters, int H, int chunk) {
    using C = CaCfg<T>;
    extern __shared__ __align__(128) uint8_t ca_smem[];
    T* sK = reinterpret_cast<T*>(ca_smem);                        // [chunk][64]
    T* sV = sK + (size_t)chunk * 64;                              // [chunk][64]
    float* sc = reinterpret_cast<float*>(sV + (size_t)chunk * 64);   // [chunk]
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ float red[CA_WARPS];
    __shared__ float opart[CA_WARPS][64];
    __shared__ int s_last;

    const int split = blockIdx.x, n_split = gridDim.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = H * 64;
    const int k0 = split * chunk;
    const int nk = min(chunk, WIPA_T_ENC - k0);                   // >= 1 by construction of chunk
    const int utt = utt_of_seq ? utt_of_seq[b] : b;
    const size_t head_off = ((size_t)utt * H + h) * WIPA_T_ENC * 64 + (size_t)k0 * 64;

    if (tid == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::mbar_init(&bar[1], 1);
        ptx::fence_barrier_init();
        const uint32_t bytes = (uint32_t)nk * 64u * (uint32_t)sizeof(T);
        ptx::mbar_arrive_expect_tx(&bar[0], bytes);
        ptx::bulk_load_1d(sK, kc + head_off, bytes, &bar[0]);
        ptx::mbar_arrive_expect_tx(&bar[1], bytes);
        ptx::bulk_load_1d(sV, vc + head_off, bytes, &bar[1]);
    }
    const int sub = tid / C::LPK;                                  // key slot of this lane group
    const int li = tid % C::LPK;                                   // 16-byte chunk of the row
    constexpr int KPI = CA_THREADS / C::LPK;                       // keys per iteration of the CTA
    float qv[C::VEC];
#pragma unroll
    for (int i = 0; i < C::VEC; ++i) qv[i] = q[(size_t)b * d + h * 64 + li * C::VEC + i];
    __syncthreads();                                               // barrier init visible to all waiters

    // ---- phase 1: scores from the K buffer ----------------------------------------------------------
    ptx::mbar_wait(&bar[0], 0);
    float mx = -INFINITY;
    for (int key = sub; key < nk; key += KPI) {
        const uint4 u = *reinterpret_cast<const uint4*>(sK + (size_t)key * 64 + li * C::VEC);
        float f[C::VEC];
        unpack16<T>(u, f);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < C::VEC; ++i) s = fmaf(f[i], qv[i], s);
#pragma unroll
        for (int off = C::LPK / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (li == 0) sc[key] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < CA_WARPS; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += CA_THREADS) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < CA_WARPS; ++w) sum += red[w];

    // ---- phase 2: P.V from the V buffer; lane owns two output dims, warps stride over keys -----------
    ptx::mbar_wait(&bar[1], 0);
    float a0 = 0.f, a1 = 0.f;
    for (int key = warp; key < nk; key += CA_WARPS) {
        const float p = sc[key];
        const float2 v2 = ld_pair(sV + (size_t)key * 64 + lane * 2);
        a0 = fmaf(p, v2.x, a0);
        a1 = fmaf(p, v2.y, a1);
    }
    opart[warp][lane * 2] = a0;
    opart[warp][lane * 2 + 1] = a1;
    __syncthreads();
    const size_t bh = (size_t)b * H + h;
    float* my_part = part + (bh * n_split + split) * CA_PART;
    if (tid < 64) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < CA_WARPS; ++w) r += opart[w][tid];
        my_part[2 + tid] = r;
        if (tid == 0) { my_part[0] = mx; my_part[1] = sum; }
    }
    // ---- ticket: the last chunk of this (seq, head) merges ---------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(&counters[bh], 1);
        s_last = (ticket == n_split - 1);
        if (s_last) counters[bh] = 0;                               // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < 64) {
        const float* pp = part + bh * n_split * CA_PART;
        float M = -INFINITY;
        for (int s = 0; s < n_split; ++s) M = fmaxf(M, __ldcg(pp + s * CA_PART));
        float L = 0.f, o = 0.f;
        for (int s = 0; s < n_split; ++s) {
            const float wgt = expf(__ldcg(pp + s * CA_PART) - M);
            L = fmaf(__ldcg(pp + s * CA_PART + 1), wgt, L);
            o = fmaf(__ldcg(pp + s * CA_PART + 2 + tid), wgt, o);
        }
        out[(size_t)b * d + h * 64 + tid] = from_f32<T>(o / L);
    }
}

// ------------------------------------------------------------------------------------------------
// persistent streamer (default): one CTA per SM, 8 consumer warps + 1 producer warp.
//
// The K/V of all (sequence, head) pairs form one long stream of UNITS: 12 chunks of CS_CK = 125 keys per pair (K chunk
// + V chunk = one ring slot, filled by two cp.async.bulk copies behind one mbarrier).  The stream is cut into
// contiguous, equally long unit ranges, one per CTA ("stream-K"): perfect balance at any batch size, and every CTA
// reads one long sequential HBM region.  A range covers whole pairs in its middle (result written directly) and at
// most two partial pairs at its ends; a partial pair leaves a flash-decoding partial (m, l, o[64], next chunk) in the
// slot of its first chunk and adds its chunk count to the pair's counter — whoever completes the count of 12 walks
// the slots in chunk order (deterministic) and writes the head's output.
//
// The producer warp runs ahead over pair boundaries, so the ring (6 x 32 KB in h16) always holds ~190 KB of loads in
// flight per SM; K/V are static inside a decode step, so under PDL the first slots are requested BEFORE
// griddepcontrol.wait and HBM does not idle between the layers' launches.  Every consumer warp owns 1/8 of each
// chunk's keys and carries a warp-local online softmax (m, l, o) across the chunks of a pair; the 8 warp partials are
// merged through shared memory at the end of the pair.  The next pair's q is prefetched one pair ahead.
// Measured on B200 at B=256: 174-181 us per launch with the consumers' math, 167 us with the math compiled out
// (-DWIPA_CA_STREAM_ONLY: ring throughput alone, 7.06 TB/s); 16 consumer warps instead of 8 were not faster.
// ------------------------------------------------------------------------------------------------
#define CS_CK 125                      // keys per chunk: 1500 = 12 x 125, every chunk is full
#define CS_NCH (WIPA_T_ENC / CS_CK)
#define CS_CONSUMERS 8
#define CS_THREADS (32 * (1 + CS_CONSUMERS))
#define CS_PART 68                     // floats per global partial: m, l, next chunk, pad, o[64]

template <typename T> struct CsCfg {
    static constexpr int CHUNK_BYTES = CS_CK * 64 * (int)sizeof(T);              // 16000 (h16) / 32000 (fp32)
    static constexpr int SLOT_BYTES = 2 * CHUNK_BYTES;
    static constexpr int STAGES = sizeof(T) == 2 ? 6 : 3;
    static constexpr int KPW = 32 / CaCfg<T>::LPK;                               // keys per warp per iteration
    static constexpr int ITERS = (CS_CK + CS_CONSUMERS * KPW - 1) / (CS_CONSUMERS * KPW);
    static constexpr int SMEM = STAGES * SLOT_BYTES + 2 * CS_CONSUMERS * CA_PART * 4 + 2 * STAGES * 8 + 64;
};

template <typename T> __device__ __forceinline__ float ca_exp(float x);
template <> __device__ __forceinline__ float ca_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ float ca_exp<h16>(float x) { return __expf(x); }

template <typename T>
__global__ void __launch_bounds__(CS_THREADS, 1)
cross_attention_stream_kernel(const float* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                              const int* __restrict__ utt_of_seq, T* __restrict__ out, float* __restrict__ part,
                              int* __restrict__ counters, int H, long long n_units, int kv_static) {
    using C = CaCfg<T>;
    using S = CsCfg<T>;
    extern __shared__ __align__(128) uint8_t cs_smem[];
    uint8_t* ring = cs_smem;
    float* wpart = reinterpret_cast<float*>(ring + S::STAGES * S::SLOT_BYTES);   // [2][8][CA_PART]
    uint64_t* full = reinterpret_cast<uint64_t*>(wpart + 2 * CS_CONSUMERS * CA_PART);
    uint64_t* empty = full + S::STAGES;
    __shared__ int s_total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = H * 64;
    const int u0 = (int)(n_units * blockIdx.x / gridDim.x);
    const int u1 = (int)(n_units * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < S::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], CS_CONSUMERS); }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == CS_CONSUMERS) {
        // ---- producer ------------------------------------------------------------------------------------------
        // inside a decode step utt_of_seq and the cached K/V were completed by kernels far upstream: no dependency wait
        if (!kv_static) pdl_wait();
        int s = 0;
        uint32_t ph = 1;                                           // first ring pass: every slot is free
        int pair = u0 / CS_NCH, ch = u0 - pair * CS_NCH;
        size_t base = 0;
        bool fresh = true;
        for (int u = u0; u < u1; ++u) {
            if (fresh) {
                const int b = pair / H, h = pair - b * H;
                const int utt = utt_of_seq ? utt_of_seq[b] : b;
                base = ((size_t)utt * H + h) * WIPA_T_ENC * 64;
                fresh = false;
            }
            ptx::mbar_wait(&empty[s], ph);
            if (ptx::elect_one()) {
                uint8_t* slot = ring + s * S::SLOT_BYTES;
                ptx::mbar_arrive_expect_tx(&full[s], S::SLOT_BYTES);
                ptx::bulk_load_1d(slot, kc + base + (size_t)ch * CS_CK * 64, S::CHUNK_BYTES, &full[s]);
                ptx::bulk_load_1d(slot + S::CHUNK_BYTES, vc + base + (size_t)ch * CS_CK * 64, S::CHUNK_BYTES, &full[s]);
            }
            __syncwarp();
            if (++s == S::STAGES) { s = 0; ph ^= 1; }
            if (++ch == CS_NCH) { ch = 0; ++pair; fresh = true; }
        }
    } else {
        // ---- consumers -----------------------------------------------------------------------------------------
        const int grp = lane / C::LPK;                             // key slot of this lane group within the warp
        const int li = lane % C::LPK;                              // 16-byte piece of the key row
        pdl_wait();                                                // q comes from the previous kernel
        pdl_launch_dependents();                                   // after the wait: at most two grids overlap
        int s = 0;
        uint32_t ph = 0;
        int parity = 0;
        const int n_pairs_total = (int)(n_units / CS_NCH);
        auto load_q = [&](int pair, float* dst) {
            const int b = pair / H, h = pair - b * H;
            const float4* qp = reinterpret_cast<const float4*>(q + (size_t)b * d + h * 64 + li * C::VEC);
#pragma unroll
            for (int i = 0; i < C::VEC / 4; ++i) {
                const float4 t = qp[i];
                dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
            }
        };
        float qn[C::VEC];
        if (u0 < u1) load_q(u0 / CS_NCH, qn);
        for (int u = u0; u < u1; parity ^= 1) {
            const int pair = u / CS_NCH;
            const int ch0 = u - pair * CS_NCH;
            const int ch1 = min(CS_NCH, ch0 + (u1 - u));
            const int b = pair / H, h = pair - b * H;
            float qv[C::VEC];
#pragma unroll
            for (int i = 0; i < C::VEC; ++i) qv[i] = qn[i];
            if (pair + 1 < n_pairs_total) load_q(pair + 1, qn);     // lands while this pair is being processed
            float m_run = -INFINITY, l_run = 0.f, a0 = 0.f, a1 = 0.f;
            for (int ch = ch0; ch < ch1; ++ch) {
                ptx::mbar_wait(&full[s], ph);
#ifdef WIPA_CA_STREAM_ONLY            /* experiment: ring throughput without the consumers' math */
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);
                if (++s == S::STAGES) { s = 0; ph ^= 1; }
                continue;
#endif
                const T* sK = reinterpret_cast<const T*>(ring + s * S::SLOT_BYTES);
                const T* sV = reinterpret_cast<const T*>(ring + s * S::SLOT_BYTES + S::CHUNK_BYTES);
                float sc[S::ITERS];
                float mw = -INFINITY;
#pragma unroll
                for (int it = 0; it < S::ITERS; ++it) {
                    const int key = (it * CS_CONSUMERS + warp) * S::KPW + grp;
                    float v = -INFINITY;
                    if (key < CS_CK) {
                        const uint4 u4 = *reinterpret_cast<const uint4*>(sK + (size_t)key * 64 + li * C::VEC);
                        float f[C::VEC];
                        unpack16<T>(u4, f);
                        v = 0.f;
#pragma unroll
                        for (int i = 0; i < C::VEC; ++i) v = fmaf(f[i], qv[i], v);
                    }
#pragma unroll
                    for (int off = C::LPK / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    sc[it] = v;                                    // -inf for absent keys (whole lane group absent)
                    mw = fmaxf(mw, v);
                }
#pragma unroll
                for (int off = C::LPK; off < 32; off <<= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, off));
                const float m_new = fmaxf(m_run, mw);              // finite: every warp owns >= 1 key of every chunk
                const float alpha = ca_exp<T>(m_run - m_new);      // 0 on the first chunk
                float lsum = 0.f;
#pragma unroll
                for (int it = 0; it < S::ITERS; ++it) {
                    sc[it] = ca_exp<T>(sc[it] - m_new);
                    lsum += sc[it];
                }
#pragma unroll
                for (int off = C::LPK; off < 32; off <<= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
                l_run = fmaf(l_run, alpha, lsum);
                a0 *= alpha; a1 *= alpha;
                m_run = m_new;
#pragma unroll
                for (int it = 0; it < S::ITERS; ++it) {
#pragma unroll
                    for (int gg = 0; gg < S::KPW; ++gg) {
                        const int key = (it * CS_CONSUMERS + warp) * S::KPW + gg;
                        const float p = __shfl_sync(0xffffffffu, sc[it], gg * C::LPK);
                        if (key < CS_CK) {
                            const float2 v2 = ld_pair(sV + (size_t)key * 64 + lane * 2);
                            a0 = fmaf(p, v2.x, a0);
                            a1 = fmaf(p, v2.y, a1);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);        // this warp is done with the slot
                if (++s == S::STAGES) { s = 0; ph ^= 1; }
            }
            u += ch1 - ch0;
            // ---- merge the 8 warp partials of this pair (or part of a pair) -------------------------------------------
            float* wp = wpart + (parity * CS_CONSUMERS + warp) * CA_PART;
            wp[2 + lane * 2] = a0;
            wp[3 + lane * 2] = a1;
            if (lane == 0) { wp[0] = m_run; wp[1] = l_run; }
            asm volatile("bar.sync 1, %0;" ::"n"(CS_CONSUMERS * 32) : "memory");
            if (warp < 2) {
                const int e = warp * 32 + lane;
                const float* pp = wpart + parity * CS_CONSUMERS * CA_PART;
                float M = pp[0];
#pragma unroll
                for (int w = 1; w < CS_CONSUMERS; ++w) M = fmaxf(M, pp[w * CA_PART]);
                float L = 0.f, o = 0.f;
#pragma unroll
                for (int w = 0; w < CS_CONSUMERS; ++w) {
                    const float wgt = ca_exp<T>(pp[w * CA_PART] - M);
                    L = fmaf(pp[w * CA_PART + 1], wgt, L);
                    o = fmaf(pp[w * CA_PART + 2 + e], wgt, o);
                }
                if (ch1 - ch0 == CS_NCH) {
                    out[(size_t)b * d + h * 64 + e] = from_f32<T>(o / L);
                } else {
                    float* gp = part + ((size_t)pair * CS_NCH + ch0) * CS_PART;
                    gp[4 + e] = o;
                    if (e == 0) { gp[0] = M; gp[1] = L; gp[2] = (float)ch1; }
                    __threadfence();
                    asm volatile("bar.sync 2, 64;" ::: "memory");    // both merging warps have published their halves
                    if (e == 0) {
                        const int before = atomicAdd(&counters[pair], ch1 - ch0);
                        s_total = before + (ch1 - ch0);
                        if (before + (ch1 - ch0) == CS_NCH) counters[pair] = 0;   // ready for the next launch
                    }
                    asm volatile("bar.sync 2, 64;" ::: "memory");
                    if (s_total == CS_NCH) {                        // all 12 chunks of this pair are in: walk the slots in order
                        __threadfence();
                        const float* g0 = part + (size_t)pair * CS_NCH * CS_PART;
                        float GM = -INFINITY;
                        for (int sl = 0; sl < CS_NCH; sl = (int)__ldcg(g0 + sl * CS_PART + 2)) GM = fmaxf(GM, __ldcg(g0 + sl * CS_PART));
                        float GL = 0.f, go = 0.f;
                        for (int sl = 0; sl < CS_NCH; sl = (int)__ldcg(g0 + sl * CS_PART + 2)) {
                            const float wgt = ca_exp<T>(__ldcg(g0 + sl * CS_PART) - GM);
                            GL = fmaf(__ldcg(g0 + sl * CS_PART + 1), wgt, GL);
                            go = fmaf(__ldcg(g0 + sl * CS_PART + 4 + e), wgt, go);
                        }
                        out[(size_t)b * d + h * 64 + e] = from_f32<T>(go / GL);
                    }
                }
            }
        }
    }
}

template <typename T>
static int launch_cross_attention_stream(const float* q, const T* k, const T* v, const int* utt_of_seq, T* out, float* part,
                                         int* counters, int Bs, int H, int kv_static, cudaStream_t st) {
    using S = CsCfg<T>;
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(cross_attention_stream_kernel<T>, (size_t)S::SMEM, attr));
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WIPA_CUDA_CHECK(cudaGetDevice(&dev));
        WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long n_units = (long long)Bs * H * CS_NCH;
    const int grid = n_units < n_sm ? (int)n_units : n_sm;
    WIPA_CUDA_CHECK(wipa_launch_c(4, cross_attention_stream_kernel<T>, dim3(grid), dim3(CS_THREADS), (size_t)S::SMEM, st, q, k, v,
                                utt_of_seq, out, part, counters, H, n_units, kv_static));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

int cross_attention_default_split(int elem_bytes, int Bs, int H) {
    // legacy kernel only: chunk sized for ~48 KB of K+V per CTA (4 CTAs / SM resident): 188 keys in h16, 94 in fp32
    (void)Bs; (void)H;
    return elem_bytes == 2 ? 8 : 16;
}

template <typename T>
int launch_cross_attention(const float* q, const T* k, const T* v, const int* utt_of_seq, T* out, float* part,
                           int* counters, int Bs, int H, int n_split, int kv_static, cudaStream_t st) {
    static const int legacy = [] { const char* e = getenv("WIPA_CA_LEGACY"); return (e && *e) ? atoi(e) : 0; }();
    if (!legacy) return launch_cross_attention_stream<T>(q, k, v, utt_of_seq, out, part, counters, Bs, H, kv_static, st);
    WIPA_CHECK(n_split >= 1 && n_split <= 64, WIPA_EINVAL, "cross_attention: n_split %d out of range", n_split);
    int chunk = cdiv(WIPA_T_ENC, n_split);
    chunk = (chunk + 3) & ~3;                                       // keeps every chunk's byte count a multiple of 16
    WIPA_CHECK((n_split - 1) * chunk < WIPA_T_ENC, WIPA_EINVAL, "cross_attention: n_split %d leaves an empty chunk", n_split);
    const size_t smem = (size_t)chunk * 64 * sizeof(T) * 2 + (size_t)chunk * sizeof(float);
    static SmemAttr attr;                              // one per template instantiation
    WIPA_TRY(wipa_ensure_smem(cross_attention_kernel<T>, smem, attr));
    dim3 grid(n_split, H, Bs);
    cross_attention_kernel<T><<<grid, CA_THREADS, smem, st>>>(q, k, v, utt_of_seq, out, part, counters, H, chunk);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
template int launch_cross_attention<float>(const float*, const float*, const float*, const int*, float*, float*, int*, int, int, int, int, cudaStream_t);
template int launch_cross_attention<h16>(const float*, const h16*, const h16*, const int*, h16*, float*, int*, int, int, int, int, cudaStream_t);
