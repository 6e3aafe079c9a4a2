// Decoder cross-attention over the encoder output itself ("latent" cross-attention), h16 path.
//
// HF computes, per decoder layer l and utterance u (HF:models/whisper/modeling_whisper.py:241-357),
//     K_l = E_u Wk_l^T,   V_l = E_u Wv_l^T + bv_l,   ctx_h = softmax_t(q_h . K_l[t, h]) V_l[:, h]
// with E_u = the encoder output [1500, d], the SAME matrix for every layer.  Streaming K_l and V_l of all layers costs
// L * 2 * 1500 * d elements per sequence per decode step (55 MB for whisper-small) and is what bounds the decode step.
// Because K and V are linear images of E, the projections can be moved to the query / output side:
//     q_h . K_l[t, h] = (Wk_l[h]^T q_h) . E_u[t]           = q'_h . E_u[t],      q'_h in R^d
//     ctx_h           = Wv_l[h] (sum_t p_h[t] E_u[t]) + bv_l[h] = Wv_l[h] c_h + bv_l[h]     (sum_t p_h[t] = 1)
// so one pass over E_u (1500 * d elements) serves all heads and both the "key" and the "value" role: half the bytes per
// layer, no per-layer cross-KV cache (14.2 GB -> 0.59 GB at 256 clips), no cross-K/V projection in the encoder.  The
// price is H times more arithmetic (every head works in R^d instead of R^64), which the legacy tensor pipe absorbs:
// S = Q' E^T and C = P E are [16 x keys x d] mma.sync products with M = 16 >= the number of heads.
//
//   q' = x (Wq_h^T Wk_h) + bq_h Wk_h      one [S, d] x [d, H*d] GEMM with weights folded at load time (ctx.cu)
//   this kernel: Q' [S, H, d] h16, E [U, 1500, d] h16  ->  C [S, H, d] h16 (normalised sum_t p_h[t] E[t])
//   out-projection: x += C (Wo_h Wv_h)^T + (bo + Wo bv)   one [S, H*d] x [H*d, d] GEMM, weights folded at load time
//
// Kernel: one CTA per SM.  The (sequence, chunk) units of the launch form one list that is cut into equal contiguous
// ranges, one per CTA (stream-K: exact balance at any number of sequences).  A sequence that lies inside one range is
// finished in registers; one that is cut leaves a partial (m, l, unnormalised C) per piece in global scratch, and the
// CTA that completes the sequence's chunk count merges the pieces in order.  Warp H is the TMA producer: chunks of KEYS keys
// x d columns as H 128B-swizzled tiles [KEYS][64] into a 2-stage ring.  Consumer warp w (0..H-1) owns columns
// [64w, 64w + 64) of d for BOTH products:
//   1. partial scores  Sp[w][16 x KEYS] = Q'[:, cols] E[keys, cols]^T    (A fragments of Q' live in registers per sequence)
//   2. named barrier; warp w sums row w (= head w) over the H partials, does the online-softmax bookkeeping for that
//      head (maximum by redux.sync) and writes p (h16) and the rescale factor alpha[w]; named barrier
//   3. C[:, cols] = alpha * C[:, cols] + P E[keys, cols]                 (accumulators [16 x 64] per warp, fp32);
//      the softmax denominator l = alpha * l + P x ones rides the same MMAs
// Rows >= H of the 16-row MMA tile are padding.  Keys beyond 1500 in the last chunk are zero-filled by TMA (3-D map,
// out-of-bounds rows) and masked to -inf before the softmax.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_xl = nullptr;

constexpr int XL_BULK_SPLIT = 4;              // bulk copies per full chunk (stage bytes are a multiple of 4 * 16)
constexpr float XL_LOG2E = 1.4426950408889634f;

template <int KEYS, int XL_STAGES>
struct XlCfg {
    static constexpr int PITCH = KEYS + 8;                      // floats per partial-score row / h16 per P row
    static size_t smem(int H) {
        return (size_t)XL_STAGES * KEYS * H * 128 + (size_t)H * H * PITCH * 4 + 16 * PITCH * 2 + 32 * 4 + 64 + 1024;
    }
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D[16 x 8] += A[16 x 16] B[16 x 8], h16 operands, fp32 accumulators
__device__ __forceinline__ void mma_h16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." WIPA_H16_MMA_SUFFIX ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void xl_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts_b16(uint32_t addr, h16 v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&v)) : "memory");
}
// warp maximum of floats in one redux.sync: IEEE floats order like sign-magnitude integers, so flip the magnitude bits of
// the negative ones, take the integer maximum and map back (the map is its own inverse; -inf stays the smallest)
__device__ __forceinline__ float warp_max_f32(float v) {
    int k = __float_as_int(v);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(k);
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// End of a (sequence, chunk range) segment: a whole sequence is normalised and stored from registers; a part of one
// leaves (m, l, unnormalised C) in this CTA's slot, and the CTA that completes the sequence's chunk count merges the
// slots in order.  Called by every consumer thread of the CTA (it contains named barriers).
template <int H>
__device__ __forceinline__ void xl_finish_segment(int s, int v, int bx, int nx, int ch0, int ch1, int n_chunks, long long n_units,
                                                  float (&acc)[8][4], float (&accl)[4], float m_run, h16* __restrict__ Cout,
                                                  float* __restrict__ part, int* __restrict__ counters, int slots_per_seq,
                                                  int* last_flag, int w, int lane) {
    constexpr int d = H * 64;
    constexpr int nthr = H * 32;
    const int g = lane >> 2, t = lane & 3;
    const bool row_lo = g < H, row_hi = g + 8 < H;
    if (ch0 == 0 && ch1 == n_chunks) {
        // the whole sequence was ours: normalise and store
        const float il_lo = row_lo ? 1.f / accl[0] : 0.f, il_hi = row_hi ? 1.f / accl[2] : 0.f;
        h16* c_lo = Cout + ((size_t)s * H + g) * d + w * 64 + 2 * t;
        h16* c_hi = Cout + ((size_t)s * H + g + 8) * d + w * 64 + 2 * t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (row_lo) *reinterpret_cast<uint32_t*>(c_lo + j * 8) = pack_h16x2(acc[j][0] * il_lo, acc[j][1] * il_lo);
            if (row_hi) *reinterpret_cast<uint32_t*>(c_hi + j * 8) = pack_h16x2(acc[j][2] * il_hi, acc[j][3] * il_hi);
        }
        return;
    }
    // a part of the sequence: leave (m, l, unnormalised C) in this CTA's slot of the sequence; whoever brings the
    // sequence's chunk count to n_chunks merges the slots in order (deterministic) and stores
    constexpr int SLOT_FLOATS = H * 1024 + 32;                 // [warp][lane][32 accumulators] + m[16] + l[16]
    const long long first_unit = (long long)v * n_chunks;
    const int cta_first = (int)(((first_unit + 1) * nx + n_units - 1) / n_units) - 1;          // CTA (group) holding chunk 0
    const int cta_last = (int)(((first_unit + n_chunks) * nx + n_units - 1) / n_units) - 1;   // CTA (group) holding the last chunk
    float* seq_part = part + (size_t)s * slots_per_seq * SLOT_FLOATS;
    {
        float* slot = seq_part + (size_t)(bx - cta_first) * SLOT_FLOATS;
        float4* dst = reinterpret_cast<float4*>(slot + (w * 32 + lane) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        if (lane == 0) slot[H * 1024 + w] = m_run;
        if (w == 0 && t == 0) {                                  // every warp holds the same l: warp 0 writes rows g / g + 8
            if (row_lo) slot[H * 1024 + 16 + g] = accl[0];
            if (row_hi) slot[H * 1024 + 16 + g + 8] = accl[2];
        }
    }
    __threadfence();
    xl_bar(nthr);
    if (threadIdx.x == 0) {
        const int mine = ch1 - ch0;
        const int old = atomicAdd(&counters[s], mine);
        const bool last = old + mine == n_chunks;
        if (last) counters[s] = 0;                               // ready for the next launch
        *last_flag = last ? 1 : 0;                             // next written after at least two more barriers
    }
    xl_bar(nthr);
    if (*last_flag == 0) return;
    __threadfence();
    {
        const int n_slots = cta_last - cta_first + 1;
        float M_lo = -INFINITY, M_hi = -INFINITY;
        for (int i = 0; i < n_slots; ++i) {
            const float* sl = seq_part + (size_t)i * SLOT_FLOATS + H * 1024;
            if (row_lo) M_lo = fmaxf(M_lo, __ldcg(sl + g));
            if (row_hi) M_hi = fmaxf(M_hi, __ldcg(sl + g + 8));
        }
        float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
        for (int i = 0; i < n_slots; ++i) {
            const float* sl = seq_part + (size_t)i * SLOT_FLOATS;
            const float f_lo = row_lo ? ex2_ftz((__ldcg(sl + H * 1024 + g) - M_lo) * XL_LOG2E) : 0.f;
            const float f_hi = row_hi ? ex2_ftz((__ldcg(sl + H * 1024 + g + 8) - M_hi) * XL_LOG2E) : 0.f;
            if (row_lo) l_lo = fmaf(f_lo, __ldcg(sl + H * 1024 + 16 + g), l_lo);
            if (row_hi) l_hi = fmaf(f_hi, __ldcg(sl + H * 1024 + 16 + g + 8), l_hi);
            const float4* src = reinterpret_cast<const float4*>(sl + (w * 32 + lane) * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = __ldcg(src + j);
                acc[j][0] = fmaf(f_lo, v.x, acc[j][0]); acc[j][1] = fmaf(f_lo, v.y, acc[j][1]);
                acc[j][2] = fmaf(f_hi, v.z, acc[j][2]); acc[j][3] = fmaf(f_hi, v.w, acc[j][3]);
            }
        }
        const float il_lo = row_lo ? 1.f / l_lo : 0.f, il_hi = row_hi ? 1.f / l_hi : 0.f;
        h16* c_lo = Cout + ((size_t)s * H + g) * d + w * 64 + 2 * t;
        h16* c_hi = Cout + ((size_t)s * H + g + 8) * d + w * 64 + 2 * t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (row_lo) *reinterpret_cast<uint32_t*>(c_lo + j * 8) = pack_h16x2(acc[j][0] * il_lo, acc[j][1] * il_lo);
            if (row_hi) *reinterpret_cast<uint32_t*>(c_hi + j * 8) = pack_h16x2(acc[j][2] * il_hi, acc[j][3] * il_hi);
        }
    }
}

// TILED: E arrives in the chunk-tiled, pre-swizzled layout (see the header of this file and lat_tile_offset in common.cuh):
// a chunk is ONE contiguous block of global memory that lands in shared memory with plain bulk copies, exactly as the
// tensor-map path would have swizzled it.  Otherwise E is row-major [U, T, d] behind a 3-D tensor map (H boxes per chunk).
template <int KEYS, int H, bool TILED, int XL_STAGES, bool BEAMS>
__global__ void __launch_bounds__((H + 1) * 32, 1)
cross_attention_latent_kernel(const __grid_constant__ CUtensorMap tmE, const h16* __restrict__ Et, const h16* __restrict__ Qp,
                              const int* __restrict__ utt_of_seq, h16* __restrict__ Cout, int S, int T,
                              float* __restrict__ part, int* __restrict__ counters, int slots_per_seq, int beams) {
    const int K = BEAMS ? beams : 1;            // a compile-time 1 in the greedy instantiation: its code is that of a plain sequence list
    using Cfg = XlCfg<KEYS, XL_STAGES>;
    constexpr int PITCH = Cfg::PITCH;
    extern __shared__ uint8_t xl_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xl_smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int d = H * 64;
    constexpr uint32_t stage_bytes = (uint32_t)KEYS * (uint32_t)H * 128u;
    uint8_t* sE = smem;                                                      // [stage][H tiles][KEYS][128 B], 128B-swizzled
    float* Sp = reinterpret_cast<float*>(sE + XL_STAGES * stage_bytes);      // [warp][head][PITCH]
    h16* Pm = reinterpret_cast<h16*>(Sp + H * H * PITCH);                  // [16][PITCH]
    float* alpha = reinterpret_cast<float*>(Pm + 16 * PITCH);                // [16] rescale of C, [16] final 1 / l
    uint64_t* full = reinterpret_cast<uint64_t*>(alpha + 32);
    uint64_t* empty = full + XL_STAGES;
    int* last_flag = reinterpret_cast<int*>(empty + XL_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (T + KEYS - 1) / KEYS;
    // stream-K: the (sequence, chunk) units form one list cut into equal contiguous ranges, one per CTA.  Beam search (K > 1):
    // the K beams of an utterance (sequences u K .. u K + K - 1) attend over the same E, so the list is over (utterance slot,
    // chunk) and a GROUP of K CTAs (blockIdx.x = K bx + beam) walks each range, one beam per CTA: the K readers of an E chunk run
    // within a few hundred cycles of each other and all but the first are served by L2 instead of HBM.
    const int bx = (int)blockIdx.x / K, beam = (int)blockIdx.x - bx * K, nx = (int)gridDim.x / K;
    const long long n_units = (long long)(S / K) * n_chunks;
    const long long u_lo = n_units * bx / nx, u_hi = n_units * (bx + 1) / nx;

    if (threadIdx.x == 0) {
        if (!TILED) ptx::prefetch_tensormap(&tmE);
        for (int s = 0; s < XL_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], (uint32_t)H); }
        ptx::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 16 * PITCH; i += blockDim.x) Pm[i] = f32_to_h16(0.f);
    if (TILED) {
        // the ragged last chunk of a sequence copies only its valid keys; the rest of the stage then holds whatever an
        // earlier chunk left there (finite values, multiplied by p = 0) - but never uninitialised shared memory
        uint4* z = reinterpret_cast<uint4*>(sE);
        for (int i = threadIdx.x; i < (int)(XL_STAGES * stage_bytes / 16); i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    if (threadIdx.x < 32) alpha[threadIdx.x] = 1.f;
    __syncthreads();

    if (warp == H) {
        // ---- producer: E is written by the encoder, long before this decode step: no dependency on the previous kernel
        int st = 0;
        uint32_t ph = 0;
        for (long long unit = u_lo; unit < u_hi;) {
            const int v = (int)(unit / n_chunks), ch0 = (int)(unit - (long long)v * n_chunks);
            const int ch1 = (u_hi - unit) < (long long)(n_chunks - ch0) ? ch0 + (int)(u_hi - unit) : n_chunks;
            unit += ch1 - ch0;
            const int u = utt_of_seq[v * K + beam];
            for (int ch = ch0; ch < ch1; ++ch) {
                ptx::mbar_wait(&empty[st], ph ^ 1);
                if (ptx::elect_one()) {
                    uint8_t* dst = sE + st * stage_bytes;
                    if (TILED) {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(Et) + ((size_t)u * n_chunks + ch) * stage_bytes;
                        const int valid = (T - ch * KEYS) < KEYS ? (T - ch * KEYS) : KEYS;
                        if (valid == KEYS) {                     // the whole chunk: contiguous, a few large bulk copies
                            constexpr uint32_t piece = stage_bytes / XL_BULK_SPLIT;
                            ptx::mbar_arrive_expect_tx(&full[st], stage_bytes);
#pragma unroll
                            for (int i = 0; i < XL_BULK_SPLIT; ++i) ptx::bulk_load_1d(dst + i * piece, src + i * piece, piece, &full[st]);
                        } else {                                 // ragged end of the sequence: the valid rows of each column tile
                            ptx::mbar_arrive_expect_tx(&full[st], (uint32_t)(H * valid * 128));
                            for (int h = 0; h < H; ++h)
                                ptx::bulk_load_1d(dst + h * (KEYS * 128), src + h * (KEYS * 128), (uint32_t)(valid * 128), &full[st]);
                        }
                    } else {
                        ptx::mbar_arrive_expect_tx(&full[st], stage_bytes);
                        for (int h = 0; h < H; ++h)
                            ptx::tma_load_3d(dst + h * (KEYS * 128), &tmE, &full[st], h * 64, ch * KEYS, u);
                    }
                }
                __syncwarp();
                if (++st == XL_STAGES) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers -----------------------------------------------------------------------------------------------------
    const int w = warp;
    const int g = lane >> 2, t = lane & 3;
    constexpr int nthr = H * 32;
    const bool row_lo = g < H, row_hi = g + 8 < H;
    pdl_wait();                                   // Q' comes from the GEMM in front of this kernel
    pdl_launch_dependents();
    const uint32_t sE_s = ptx::smem_u32(sE), Pm_s = ptx::smem_u32(Pm), Sp_s = ptx::smem_u32(Sp);
    int st = 0;
    uint32_t ph = 0;
    for (long long unit = u_lo; unit < u_hi;) {
        const int v = (int)(unit / n_chunks), ch0 = (int)(unit - (long long)v * n_chunks);
        const int ch1 = (u_hi - unit) < (long long)(n_chunks - ch0) ? ch0 + (int)(u_hi - unit) : n_chunks;
        unit += ch1 - ch0;
        const int s = v * K + beam;
        // A fragments of Q': rows g / g + 8 (heads), this warp's 64 columns = 4 k-steps of 16
        uint32_t qa[4][4];
        {
            const uint32_t* q_lo = reinterpret_cast<const uint32_t*>(Qp + ((size_t)s * H + g) * d + w * 64 + 2 * t);
            const uint32_t* q_hi = reinterpret_cast<const uint32_t*>(Qp + ((size_t)s * H + g + 8) * d + w * 64 + 2 * t);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qa[ks][0] = row_lo ? q_lo[ks * 8] : 0u;           // 16 h16 = 8 words per k-step
                qa[ks][1] = row_hi ? q_hi[ks * 8] : 0u;
                qa[ks][2] = row_lo ? q_lo[ks * 8 + 4] : 0u;       // columns + 8
                qa[ks][3] = row_hi ? q_hi[ks * 8 + 4] : 0u;
            }
        }
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
        float m_run = -INFINITY;                    // running maximum of head w (replicated over the lanes of warp w)
        // l = sum_t p[t] rides the tensor pipe: P x ones accumulates it in every warp with exactly the h16 weights (and the
        // rescaling) that C sees - no cross-lane sum, no cross-warp exchange.  accl[0] / accl[2]: rows g / g + 8
        float accl[4] = {0.f, 0.f, 0.f, 0.f};

        for (int ch = ch0; ch < ch1; ++ch) {
            ptx::mbar_wait(&full[st], ph);
            const uint32_t tile = sE_s + (uint32_t)st * stage_bytes + (uint32_t)w * (KEYS * 128);
            // 1. partial scores over this warp's columns: every fragment load first, then k-step-major MMAs so that
            //    consecutive instructions hit different accumulators
            {
                uint32_t bfr[KEYS / 8][2][4];
#pragma unroll
                for (int nt = 0; nt < KEYS / 8; ++nt) {
                    const int key = nt * 8 + (lane & 7);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c16 = half * 4 + (lane >> 3);
                        ldmatrix_x4(tile + (uint32_t)key * 128u + (uint32_t)((c16 ^ (key & 7)) << 4), bfr[nt][half][0], bfr[nt][half][1],
                                    bfr[nt][half][2], bfr[nt][half][3]);
                    }
                }
                float sc[KEYS / 8][4];
#pragma unroll
                for (int nt = 0; nt < KEYS / 8; ++nt) { sc[nt][0] = 0.f; sc[nt][1] = 0.f; sc[nt][2] = 0.f; sc[nt][3] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                    for (int nt = 0; nt < KEYS / 8; ++nt)
                        mma_h16(sc[nt], qa[ks], bfr[nt][ks >> 1][(ks & 1) * 2], bfr[nt][ks >> 1][(ks & 1) * 2 + 1]);
                }
#pragma unroll
                for (int nt = 0; nt < KEYS / 8; ++nt) {
                    if (row_lo) sts_f32x2(Sp_s + (uint32_t)((w * H + g) * PITCH + nt * 8 + 2 * t) * 4u, sc[nt][0], sc[nt][1]);
                    if (row_hi) sts_f32x2(Sp_s + (uint32_t)((w * H + g + 8) * PITCH + nt * 8 + 2 * t) * 4u, sc[nt][2], sc[nt][3]);
                }
            }
            xl_bar(nthr);
            // fragments of E for step 3 do not depend on the softmax: fetch them now, behind the reduction
            uint32_t vfr[KEYS / 16][4][4];
#pragma unroll
            for (int ks = 0; ks < KEYS / 16; ++ks) {
                const int key = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    const int c16 = jp * 2 + (lane >> 4);
                    ldmatrix_x4_trans(tile + (uint32_t)key * 128u + (uint32_t)((c16 ^ (key & 7)) << 4), vfr[ks][jp][0], vfr[ks][jp][1],
                                      vfr[ks][jp][2], vfr[ks][jp][3]);
                }
            }
            // 2. head w: sum the partials (two interleaved chains), maximum by redux.sync, p in h16, rescale factor
            {
                constexpr bool kTwo = KEYS > 32;
                const bool has1 = kTwo && lane < KEYS - 32;
                const uint32_t row = Sp_s + (uint32_t)(w * PITCH + lane) * 4u;
                float v0a = 0.f, v0b = 0.f, v1a = 0.f, v1b = 0.f;
#pragma unroll
                for (int ww = 0; ww < H; ww += 2) {
                    v0a += lds_f32(row + (uint32_t)(ww * H * PITCH) * 4u);
                    v0b += lds_f32(row + (uint32_t)((ww + 1) * H * PITCH) * 4u);
                    if (kTwo) {
                        v1a += lds_f32(row + (uint32_t)(ww * H * PITCH + (has1 ? 32 : 0)) * 4u);
                        v1b += lds_f32(row + (uint32_t)((ww + 1) * H * PITCH + (has1 ? 32 : 0)) * 4u);
                    }
                }
                float v0 = v0a + v0b, v1 = v1a + v1b;
                const int key0 = ch * KEYS + lane;
                if (key0 >= T) v0 = -INFINITY;
                if (!has1 || key0 + 32 >= T) v1 = -INFINITY;
                const float m_new = fmaxf(m_run, warp_max_f32(fmaxf(v0, v1)));      // finite: every chunk holds a valid key
                const float mb = m_new * XL_LOG2E;
                const float p0 = ex2_ftz(fmaf(v0, XL_LOG2E, -mb)), p1 = ex2_ftz(fmaf(v1, XL_LOG2E, -mb));
                const float a = ex2_ftz(fmaf(m_run, XL_LOG2E, -mb));
                m_run = m_new;
                sts_b16(Pm_s + (uint32_t)(w * PITCH + lane) * 2u, f32_to_h16(p0));
                if (has1) sts_b16(Pm_s + (uint32_t)(w * PITCH + 32 + lane) * 2u, f32_to_h16(p1));
                if (lane == 0) alpha[w] = a;
            }
            xl_bar(nthr);
            // 3. C = alpha * C + P E over this warp's columns
            {
                const float a_lo = alpha[g], a_hi = alpha[g + 8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[j][0] *= a_lo; acc[j][1] *= a_lo; acc[j][2] *= a_hi; acc[j][3] *= a_hi; }
                accl[0] *= a_lo; accl[1] *= a_lo; accl[2] *= a_hi; accl[3] *= a_hi;
                uint32_t pa[KEYS / 16][4];
#pragma unroll
                for (int ks = 0; ks < KEYS / 16; ++ks) {
                    const uint32_t p_lo = Pm_s + (uint32_t)(g * PITCH + ks * 16 + 2 * t) * 2u;
                    const uint32_t p_hi = p_lo + 8u * PITCH * 2u;
                    pa[ks][0] = lds32(p_lo); pa[ks][1] = lds32(p_hi); pa[ks][2] = lds32(p_lo + 16u); pa[ks][3] = lds32(p_hi + 16u);
                }
#pragma unroll
                for (int ks = 0; ks < KEYS / 16; ++ks) {
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        mma_h16(acc[2 * jp], pa[ks], vfr[ks][jp][0], vfr[ks][jp][1]);
                        mma_h16(acc[2 * jp + 1], pa[ks], vfr[ks][jp][2], vfr[ks][jp][3]);
                    }
                    mma_h16(accl, pa[ks], WIPA_H16_ONE_X2, WIPA_H16_ONE_X2);       // x ones (h16 1.0 pairs): row sums of P
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[st]);
            if (++st == XL_STAGES) { st = 0; ph ^= 1; }
        }
        xl_finish_segment<H>(s, v, bx, nx, ch0, ch1, n_chunks, n_units, acc, accl, m_run, Cout, part, counters, slots_per_seq, last_flag, w, lane);
    }
}

int xl_make_map(CUtensorMap* map, const void* E, int U, int T, int d, int keys) {
    cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)T, (cuuint64_t)U};
    cuuint64_t strides[2] = {(cuuint64_t)d * 2, (cuuint64_t)T * d * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)keys, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_xl(map, WIPA_H16_TMA_TYPE, 3, const_cast<void*>(E), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wipa_set_error("cross_attention_latent: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return WIPA_ECUDA;
    }
    return WIPA_OK;
}

template <int KEYS, int H, bool TILED, int XL_STAGES = 2, bool BEAMS = false>
int xl_launch(const CUtensorMap& tm, const h16* Et, const h16* Qp, const int* utt_of_seq, h16* C, int S, int T, int n_sm, float* part,
              size_t part_floats, int* counters, int K, cudaStream_t st) {
    const size_t smem = XlCfg<KEYS, XL_STAGES>::smem(H);
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(cross_attention_latent_kernel<KEYS, H, TILED, XL_STAGES, BEAMS>, smem, attr));
    if (!BEAMS) K = 1;
    const int n_chunks = cdiv(T, KEYS);
    const long long n_units = (long long)(S / K) * n_chunks;      // per group of K CTAs (K = 1: per CTA)
    const int groups = n_units < n_sm / K ? (int)n_units : n_sm / K;
    const int grid = groups * K;
    // a sequence is cut by at most n_chunks / (shortest range) range boundaries
    const int slots_per_seq = n_chunks / (int)(n_units / groups) + 2;
    WIPA_CHECK((size_t)S * slots_per_seq * (H * 1024 + 32) <= part_floats, WIPA_EINVAL,
               "cross_attention_latent: partial scratch too small for %d sequences", S);
    WIPA_CUDA_CHECK(wipa_launch_c(4, cross_attention_latent_kernel<KEYS, H, TILED, XL_STAGES, BEAMS>, dim3(grid), dim3((H + 1) * 32), smem, st, tm, Et, Qp,
                                  utt_of_seq, C, S, T, part, counters, slots_per_seq, K));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

}  // namespace

// one instantiation per Whisper width below large (heads = d / 64): tiny 6, base 8, small 12, medium 16
int cross_attention_latent_supported(int H) { return H == 6 || H == 8 || H == 12 || H == 16; }

// floats of partial scratch that any launch with S <= max_seqs sequences can need on a device with n_sm SMs
size_t cross_attention_latent_scratch_floats(int H, int max_seqs, int n_sm) {
    return (size_t)(2 * n_sm + 3 * max_seqs + 64) * (size_t)(H * 1024 + 32);
}

// Chunk size / ring depth: the stages of keys x d x 2 bytes plus H x H partial-score rows must fit 227 KB of shared memory.
// Default up to 12 heads: 32 keys x 3 stages - measured at small / 256 sequences per decode step: 2304 us against 2383 for 48 keys
// x 2 stages, 2345 for 32 x 4 and 2433 for 32 x 2 (the depth of the prefetch ring matters more than the chunk size; 48 x 3 does
// not fit).  16 heads: 32 keys x 2 stages.  WIPA_XL_KEYS=48 / WIPA_XL_STAGES=2|3|4 select the others.  Read once per process:
// the tiled layout of the encoder output depends on the chunk size.
static int xl_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
int cross_attention_latent_keys(int H) {
    static const int env_keys = xl_env("WIPA_XL_KEYS", 32);
    return (H <= 12 && env_keys == 48) ? 48 : 32;
}
static int xl_stages(int H) {
    static const int env_stages = xl_env("WIPA_XL_STAGES", 0);
    if (H > 12 || cross_attention_latent_keys(H) == 48) return 2;
    return env_stages == 4 ? 4 : (env_stages == 2 ? 2 : 3);
}

// elements of the chunk-tiled image of one utterance's encoder output: whole chunks of `keys` keys (the tail is zero padding)
size_t cross_attention_latent_tiled_elems(int H, int T) {
    const int keys = cross_attention_latent_keys(H);
    return (size_t)cdiv(T, keys) * keys * H * 64;
}

// Qp: h16 [S, H, d] absorbed queries; E: encoder output (d = 64 H) as h16 [U, T, d] (tiled = 0) or in the chunk-tiled layout
// (tiled = 1: [U][chunk][h][key][64 swizzled], common.cuh lat_tile_offset); utt_of_seq: int [S]; C: h16 [S, H, d];
// part / counters: partial scratch (cross_attention_latent_scratch_floats) and int [S] zeroed once (self-resetting)
int launch_cross_attention_latent(const h16* Qp, const h16* E, int tiled, int U, const int* utt_of_seq, h16* C, int S, int H, int T,
                                  float* part, size_t part_floats, int* counters, cudaStream_t st, int beams) {
    WIPA_CHECK(cross_attention_latent_supported(H), WIPA_EUNSUPPORTED, "cross_attention_latent: %d heads (6, 8, 12 or 16)", H);
    WIPA_CHECK(S >= 1 && U >= 1 && T >= 1 && part && counters, WIPA_EINVAL, "cross_attention_latent: bad argument");
    // beams > 1: sequences u * beams .. u * beams + beams - 1 share utt_of_seq (the beams of one utterance); they are walked by
    // a group of `beams` CTAs.  Anything else (or more beams than a tenth of the SMs) falls back to one list of sequences.
    const int K = (beams >= 2 && beams <= 8 && S % beams == 0) ? beams : 1;
    if (g_encode_xl == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WIPA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WIPA_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, WIPA_ECUDA, "cuTensorMapEncodeTiled not available");
        g_encode_xl = reinterpret_cast<EncodeTiledFn>(fn);
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WIPA_CUDA_CHECK(cudaGetDevice(&dev));
        WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int keys = cross_attention_latent_keys(H);
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (tiled && keys == 32 && H <= 12 && K > 1 && xl_stages(H) == 3) {
        // beam search on the default layout: groups of K CTAs per key range (other layouts walk one list of sequences)
        if (H == 6) return xl_launch<32, 6, true, 3, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        if (H == 8) return xl_launch<32, 8, true, 3, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        return xl_launch<32, 12, true, 3, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
    }
    if (tiled && keys == 32 && H <= 12) {
#define XL_CASE(HH, ST) if (H == HH && xl_stages(H) == ST) return xl_launch<32, HH, true, ST>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st)
        XL_CASE(6, 2); XL_CASE(6, 3); XL_CASE(6, 4); XL_CASE(8, 2); XL_CASE(8, 3); XL_CASE(8, 4); XL_CASE(12, 2); XL_CASE(12, 3); XL_CASE(12, 4);
#undef XL_CASE
    }
    if (tiled) {
        switch (H) {
            case 6: return xl_launch<48, 6, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
            case 8: return xl_launch<48, 8, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
            case 12: return xl_launch<48, 12, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
            default: return xl_launch<32, 16, true>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        }
    }
    WIPA_TRY(xl_make_map(&tm, E, U, T, H * 64, H <= 12 ? 48 : 32));       // the row-major path keeps 48 keys x 2 stages up to 12 heads
    switch (H) {
        case 6: return xl_launch<48, 6, false>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        case 8: return xl_launch<48, 8, false>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        case 12: return xl_launch<48, 12, false>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
        default: return xl_launch<32, 16, false>(tm, E, Qp, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
    }
}
