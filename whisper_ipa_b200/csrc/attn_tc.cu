// Encoder self-attention on the 5th-generation tensor cores (h16 path).
//
// Non-causal attention over T = 1500 frames, head_dim 64 (HF:models/whisper/modeling_whisper.py:284-357; the 64^-0.5
// scaling is folded into the q projection).  One CTA owns TWO 128-query tiles of one (clip, head) and walks the keys
// in blocks of 128; the two tiles share every K / V block (half the L2 traffic) and keep the MUFU pipe busy in turns:
//
//   warp 0      TMA producer: both Q tiles once, then K / V blocks ([128 keys][64] each, 128B-swizzled), 2-deep ring
//   warp 1      tcgen05 issuer, per key block j and tile g:
//                   S_g   = Q_g K_j^T        (M128 x N128 x K64, fp32 in TMEM, one buffer per tile)
//                   O_g,j = P_g,j V_j        (M128 x N64 x K128; P is the h16 tile the softmax warps left in shared
//                                             memory, V_j is consumed as an MN-major B operand straight from its TMA
//                                             image - no transpose anywhere; double-buffered in TMEM)
//   warps 2-5   softmax of tile 0, warps 6-9 softmax of tile 1, one query row per thread: a max pass and an exp pass
//               over S in TMEM (64 columns in registers at a time), running max / sum in fp32 (exp2 with log2e folded
//               into one FFMA), P -> h16 -> swizzled shared memory, then fold the PREVIOUS block's O_{j-1} from TMEM
//               into the fp32 register accumulator (o = o * alpha + O_{j-1}).  The tensor pipe therefore never waits
//               for a rescale: every P V product starts from zero in its own TMEM buffer.
//
// Tail handling: key indices >= T get -inf scores (their K/V rows are zero-filled by TMA), query rows >= T are
// computed on zero-filled Q and never stored.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int FA_BQ = 256;                                    // two 128-row tiles per CTA
constexpr int FA_BK = 128;
constexpr int FA_STAGES = 2;
constexpr int FA_TILE_BYTES = 128 * 64 * 2;                   // one [128][64] h16 tile
constexpr int FA_P_BYTES = 2 * FA_TILE_BYTES;                 // [128 q][128 keys] as two K-chunks of 64 keys
constexpr int FA_SMEM = FA_TILE_BYTES * (2 + 2 * FA_STAGES) + 4 * FA_P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int FA_TMEM_COLS = 512;                             // S: 2 tiles x 128 columns, O: 2 tiles x 2 x 64 columns
constexpr int FA_THREADS = 64 + 256;
constexpr float LOG2E = 1.4426950408889634f;
#ifndef FA_PF_AHEAD
#define FA_PF_AHEAD 2
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_fa = nullptr;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// V_j is consumed as an MN-major B operand with the 128-byte swizzle: rows are K indices (keys), 128 bytes = 64 N
// elements per row, 8-row groups 1024 bytes apart (cute/atom/mma_traits_sm100.hpp: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)),
// m = 1 here), i.e. exactly the TMA image of a [128 keys][64] tile; 16 keys per MMA = 2048 bytes per K step.
__host__ __device__ constexpr uint32_t idesc_h16(int M, int N, int b_mn_major) {
    return (1u << 4) | (WIPA_H16_IDESC_FMT << 7) | (WIPA_H16_IDESC_FMT << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(FA_THREADS, 1)
enc_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, h16* __restrict__ out, int H, int T) {
    extern __shared__ uint8_t fa_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                        // [2 tiles]
    uint8_t* sK = sQ + 2 * FA_TILE_BYTES;                      // [stage]
    uint8_t* sV = sK + FA_STAGES * FA_TILE_BYTES;              // [stage]
    uint8_t* sP = sV + FA_STAGES * FA_TILE_BYTES;              // [tile][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * FA_P_BYTES);
    uint64_t* q_full = bars;                                   // 1
    uint64_t* kv_full = bars + 1;                              // FA_STAGES
    uint64_t* kv_empty = kv_full + FA_STAGES;                  // FA_STAGES
    uint64_t* s_full = kv_empty + FA_STAGES;                   // [tile]
    uint64_t* p_ready = s_full + 2;                            // [tile]
    uint64_t* o_full = p_ready + 2;                            // [tile][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BQ;
    const int bh = blockIdx.y;
    const int nb = (T + FA_BK - 1) / FA_BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmK);
        ptx::prefetch_tensormap(&tmV);
    }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(q_full, 1);
            for (int s = 0; s < FA_STAGES; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
            for (int g = 0; g < 2; ++g) { ptx::mbar_init(&s_full[g], 1); ptx::mbar_init(&p_ready[g], 128); }
            for (int i = 0; i < 4; ++i) ptx::mbar_init(&o_full[i], 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, FA_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;                         // + 128 * tile
    const uint32_t tmem_O = tmem_base + 256;                   // + 128 * tile + 64 * buf

    // role warps run warp-uniformly and elect one lane per issue (bare UTMALDG / UTCHMMA in SASS, descriptors in
    // uniform registers)
    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(q_full, 2 * FA_TILE_BYTES);
            ptx::tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
            ptx::tma_load_3d(sQ + FA_TILE_BYTES, &tmQ, q_full, 0, q0 + 128, bh);
        }
        __syncwarp();
        for (int j = 0; j < nb; ++j) {
            const int s = j & 1;
            if (j >= FA_STAGES) ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                ptx::tma_load_3d(sK + s * FA_TILE_BYTES, &tmK, &kv_full[s], 0, j * FA_BK, bh);
                ptx::tma_load_3d(sV + s * FA_TILE_BYTES, &tmV, &kv_full[s], 0, j * FA_BK, bh);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = idesc_h16(128, 128, 0);
        constexpr uint32_t idesc_o = idesc_h16(128, 64, 1);
        const uint32_t q_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sK));
        const uint32_t p_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sP));
        const uint32_t v_lo0 = ((ptx::smem_u32(sV) & 0x3FFFFu) >> 4) | ((1024u >> 4) << 16);     // MN-major: LBO field = 1024 B
        ptx::mbar_wait(q_full, 0);
        for (int j = 0; j <= nb; ++j) {                           // iteration nb only drains the last P V products
            if (j < nb) {
                ptx::mbar_wait(&kv_full[j & 1], (j >> 1) & 1);
                ptx::tc_fence_after();
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (j > 0) {                                      // P_g(j-1) is in shared memory and S_g has been drained
                    ptx::mbar_wait(&p_ready[g], (j - 1) & 1);
                    ptx::tc_fence_after();
                }
                if (ptx::elect_one()) {
                    if (j < nb) {                                 // S_g = Q_g K_j^T
                        const uint32_t q_lo = q_lo0 + (uint32_t)g * (FA_TILE_BYTES >> 4);
                        const uint32_t k_lo = k_lo0 + (uint32_t)(j & 1) * (FA_TILE_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_h16(tmem_S + (uint32_t)g * 128u, ptx::smem_desc_sw128(q_lo + 2 * k),
                                           ptx::smem_desc_sw128(k_lo + 2 * k), idesc_s, k != 0 ? 1u : 0u);
                        ptx::umma_commit(&s_full[g]);
                    }
                    if (j > 0) {                                  // O_g(j-1) = P_g(j-1) V_{j-1}, 16 keys per MMA
                        const int jp = j - 1;
                        const uint32_t p_lo = p_lo0 + (uint32_t)(g * 2 + (jp & 1)) * (FA_P_BYTES >> 4);
                        const uint32_t v_lo = v_lo0 + (uint32_t)(jp & 1) * (FA_TILE_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            ptx::umma_h16(tmem_O + (uint32_t)g * 128u + (uint32_t)(jp & 1) * 64u,
                                           ptx::smem_desc_sw128(p_lo + (uint32_t)(k >> 2) * (FA_TILE_BYTES >> 4) + (uint32_t)(k & 3) * 2u),
                                           ptx::smem_desc_sw128(v_lo + (uint32_t)k * (2048u >> 4)), idesc_o, k != 0 ? 1u : 0u);
                        ptx::umma_commit(&o_full[g * 2 + (jp & 1)]);
                        if (g == 1) ptx::umma_commit(&kv_empty[jp & 1]);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        const int g = (warp - 2) >> 2;                            // which query tile
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may touch
        const int r = quarter * 32 + lane;                        // query row of this thread within the tile
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t ts = tmem_S + (uint32_t)g * 128u + lane_off;
        const uint32_t to = tmem_O + (uint32_t)g * 128u + lane_off;
        uint64_t* my_s_full = &s_full[g];
        uint64_t* my_p_ready = &p_ready[g];
        uint64_t* my_o_full = &o_full[g * 2];
        uint8_t* myP = sP + g * 2 * FA_P_BYTES;
        float o[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;
        float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
        const uint32_t p_row = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);

        auto fold_o = [&](int j, float alpha) {                   // o = o * alpha + O_j
            ptx::mbar_wait(&my_o_full[j & 1], (j >> 1) & 1);
            ptx::tc_fence_after();
            float t[64];
            ptx::tmem_ld32(to + (uint32_t)(j & 1) * 64u, t);
            ptx::tmem_ld32(to + (uint32_t)(j & 1) * 64u + 32u, t + 32);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = fmaf(o[i], alpha, t[i]);
        };

        for (int j = 0; j < nb; ++j) {
            ptx::mbar_wait(my_s_full, j & 1);
            ptx::tc_fence_after();
            const int kbase = j * FA_BK;
            const bool tail = kbase + FA_BK > T;
            float s[64];
            // ---- pass 1: row max over the 128 scores ------------------------------------------------------------
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                ptx::tmem_ld32(ts + hf * 64, s);
                ptx::tmem_ld32(ts + hf * 64 + 32, s + 32);
                ptx::tmem_ld_wait();
                if (tail) {
#pragma unroll
                    for (int c = 0; c < 64; ++c) if (kbase + hf * 64 + c >= T) s[c] = -INFINITY;
                }
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    mx0 = fmaxf(mx0, s[c]); mx1 = fmaxf(mx1, s[c + 1]); mx2 = fmaxf(mx2, s[c + 2]); mx3 = fmaxf(mx3, s[c + 3]);
                }
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const float m_new = fmaxf(m, mx);
            const float alpha = ex2((m - m_new) * LOG2E);         // 0 on the first block (m = -inf)
            const float nm = -m_new * LOG2E;
            m = m_new;
            // ---- pass 2: p = exp(s - m), h16 P tile into swizzled shared memory ---------------------------------------
            float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
            const uint32_t prow_s = ptx::smem_u32(myP + (j & 1) * FA_P_BYTES + p_row);
#pragma unroll
            for (int hf = 1; hf >= 0; --hf) {                      // second half first: it is already in registers
                if (hf == 0) {
                    ptx::tmem_ld32(ts, s);
                    ptx::tmem_ld32(ts + 32, s + 32);
                    ptx::tmem_ld_wait();
                    if (tail) {
#pragma unroll
                        for (int c = 0; c < 64; ++c) if (kbase + c >= T) s[c] = -INFINITY;
                    }
                }
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {                   // 8 keys = one 16-byte piece of the swizzled row
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[i] = ex2(fmaf(s[c8 * 8 + i], LOG2E, nm));
                    sum0 += p[0] + p[4]; sum1 += p[1] + p[5]; sum2 += p[2] + p[6]; sum3 += p[3] + p[7];
                    uint4 u;
                    u.x = pack_h16x2(p[0], p[1]); u.y = pack_h16x2(p[2], p[3]);
                    u.z = pack_h16x2(p[4], p[5]); u.w = pack_h16x2(p[6], p[7]);
                    ptx::sts128(prow_s + hf * FA_TILE_BYTES + (((uint32_t)c8 ^ sw) << 4), u);
                }
            }
            l = fmaf(l, alpha, (sum0 + sum1) + (sum2 + sum3));
            ptx::fence_proxy_async();                             // P visible to the tensor core's async proxy
            ptx::tc_fence_before();                               // our TMEM reads of S are complete
            ptx::mbar_arrive(my_p_ready);
            if (j > 0) fold_o(j - 1, alpha_prev);
            alpha_prev = alpha;
        }
        fold_o(nb - 1, alpha_prev);
        const int t = q0 + g * 128 + r;
        if (t < T) {
            const float inv = 1.0f / l;
            const int b = bh / H, h = bh - b * H;
            h16* dst = out + ((size_t)b * T + t) * ((size_t)H * 64) + (size_t)h * 64;
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
                uint4 u;
                u.x = pack_h16x2(o[i] * inv, o[i + 1] * inv); u.y = pack_h16x2(o[i + 2] * inv, o[i + 3] * inv);
                u.z = pack_h16x2(o[i + 4] * inv, o[i + 5] * inv); u.w = pack_h16x2(o[i + 6] * inv, o[i + 7] * inv);
                *reinterpret_cast<uint4*>(dst + i) = u;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

// ---- persistent form ----------------------------------------------------------------------------------------------------
// Measured on the kernel above (scripts/fa_probe.py, T = 256 .. 3072): a CTA costs 6.6 us + 1.53 us per 128-key block, i.e. a
// quarter of a whisper CTA (12 blocks, 25 us) is prologue and epilogue that nothing overlaps - CTA launch, barrier / TMEM set-up,
// the first Q / K / V round trip, the ramp of the S -> softmax -> P V pipeline, the last fold, the stores, the teardown - and
// with one CTA per SM (224 KB of shared memory, all 512 TMEM columns) the next CTA cannot start under it.  Here one CTA per SM
// walks the (query tile pair, clip x head) items i = blockIdx.x, blockIdx.x + gridDim.x, ... as ONE stream of key blocks: the
// barriers, rings and TMEM stay alive, the producer refills Q as soon as the last S products of an item have retired, the K / V
// ring runs ahead into the next item, and the issuer starts the next item's first S products while the softmax warps are
// still folding and storing the current one.
__global__ void __launch_bounds__(FA_THREADS, 1)
enc_attention_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                   const __grid_constant__ CUtensorMap tmV, h16* __restrict__ out, int H, int T, int n_items) {
    extern __shared__ uint8_t fa_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                        // [2 tiles]
    uint8_t* sK = sQ + 2 * FA_TILE_BYTES;                      // [stage]
    uint8_t* sV = sK + FA_STAGES * FA_TILE_BYTES;              // [stage]
    uint8_t* sP = sV + FA_STAGES * FA_TILE_BYTES;              // [tile][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * FA_P_BYTES);
    uint64_t* q_full = bars;                                   // 1
    uint64_t* q_empty = bars + 1;                              // 1
    uint64_t* kv_full = bars + 2;                              // FA_STAGES
    uint64_t* kv_empty = kv_full + FA_STAGES;                  // FA_STAGES
    uint64_t* s_full = kv_empty + FA_STAGES;                   // [tile]
    uint64_t* p_ready = s_full + 2;                            // [tile]
    uint64_t* o_full = p_ready + 2;                            // [tile][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = (T + FA_BK - 1) / FA_BK;
    const int n_qt = (T + FA_BQ - 1) / FA_BQ;                  // items are ordered query tile fastest: the tiles of one (clip, head)
                                                               // run on neighbouring CTAs at the same time and share K / V in L2
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmK);
        ptx::prefetch_tensormap(&tmV);
    }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(q_full, 1);
            ptx::mbar_init(q_empty, 1);
            for (int s = 0; s < FA_STAGES; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
            for (int g = 0; g < 2; ++g) { ptx::mbar_init(&s_full[g], 1); ptx::mbar_init(&p_ready[g], 128); }
            for (int i = 0; i < 4; ++i) ptx::mbar_init(&o_full[i], 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, FA_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;                         // + 128 * tile
    const uint32_t tmem_O = tmem_base + 256;                   // + 128 * tile + 64 * buf

    if (warp == 0) {
        int c = 0;                                                // running key-block index of this CTA (ring slot and phase)
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int bh = item / n_qt, q0 = (item - bh * n_qt) * FA_BQ;
            if (it > 0) ptx::mbar_wait(q_empty, (it - 1) & 1);    // the previous item's last S products have retired
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(q_full, 2 * FA_TILE_BYTES);
                ptx::tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
                ptx::tma_load_3d(sQ + FA_TILE_BYTES, &tmQ, q_full, 0, q0 + 128, bh);
                // the next item's Q tiles and first two K / V blocks go to L2 now: their loads are issued a block ahead of their
                // use, which hides an L2 hit but not an HBM round trip
                const int nxt = item + (int)gridDim.x;
                if (nxt < n_items) {
                    const int bh2 = nxt / n_qt, q2 = (nxt - bh2 * n_qt) * FA_BQ;
                    ptx::tma_prefetch_3d(&tmQ, 0, q2, bh2);
                    ptx::tma_prefetch_3d(&tmQ, 0, q2 + 128, bh2);
                    ptx::tma_prefetch_3d(&tmK, 0, 0, bh2);
                    ptx::tma_prefetch_3d(&tmV, 0, 0, bh2);
                    if (nb > 1) {
                        ptx::tma_prefetch_3d(&tmK, 0, FA_BK, bh2);
                        ptx::tma_prefetch_3d(&tmV, 0, FA_BK, bh2);
                    }
                }
            }
            __syncwarp();
            for (int j = 0; j < nb; ++j, ++c) {
                const int s = c & 1;
                if (c >= FA_STAGES) ptx::mbar_wait(&kv_empty[s], ((c >> 1) & 1) ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                    ptx::tma_load_3d(sK + s * FA_TILE_BYTES, &tmK, &kv_full[s], 0, j * FA_BK, bh);
                    ptx::tma_load_3d(sV + s * FA_TILE_BYTES, &tmV, &kv_full[s], 0, j * FA_BK, bh);
                    if (j + FA_PF_AHEAD < nb) {                   // the ring is two deep: keep L2 a few blocks ahead of it
                        ptx::tma_prefetch_3d(&tmK, 0, (j + FA_PF_AHEAD) * FA_BK, bh);
                        ptx::tma_prefetch_3d(&tmV, 0, (j + FA_PF_AHEAD) * FA_BK, bh);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = idesc_h16(128, 128, 0);
        constexpr uint32_t idesc_o = idesc_h16(128, 64, 1);
        const uint32_t q_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sK));
        const uint32_t p_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sP));
        const uint32_t v_lo0 = ((ptx::smem_u32(sV) & 0x3FFFFu) >> 4) | ((1024u >> 4) << 16);     // MN-major: LBO field = 1024 B
        const int my_items = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total = my_items * nb;                          // key blocks of this CTA's whole stream
        // the stream of blocks c = 0 .. total - 1 (item c / nb, block c % nb); iteration c issues S(c) and P V(c - 1)
        auto issue_s = [&](int c, int g, bool last_of_item) {    // S_g = Q_g K^T (one elected lane)
            const uint32_t q_lo = q_lo0 + (uint32_t)g * (FA_TILE_BYTES >> 4);
            const uint32_t k_lo = k_lo0 + (uint32_t)(c & 1) * (FA_TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ptx::umma_h16(tmem_S + (uint32_t)g * 128u, ptx::smem_desc_sw128(q_lo + 2 * k), ptx::smem_desc_sw128(k_lo + 2 * k), idesc_s,
                               k != 0 ? 1u : 0u);
            ptx::umma_commit(&s_full[g]);
            if (g == 1 && last_of_item) ptx::umma_commit(q_empty);            // Q is through once these retire
        };
        auto issue_pv = [&](int cp, int g) {                      // O_g(cp) = P_g(cp) V, 16 keys per MMA
            const uint32_t p_lo = p_lo0 + (uint32_t)(g * 2 + (cp & 1)) * (FA_P_BYTES >> 4);
            const uint32_t v_lo = v_lo0 + (uint32_t)(cp & 1) * (FA_TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                ptx::umma_h16(tmem_O + (uint32_t)g * 128u + (uint32_t)(cp & 1) * 64u,
                               ptx::smem_desc_sw128(p_lo + (uint32_t)(k >> 2) * (FA_TILE_BYTES >> 4) + (uint32_t)(k & 3) * 2u),
                               ptx::smem_desc_sw128(v_lo + (uint32_t)k * (2048u >> 4)), idesc_o, k != 0 ? 1u : 0u);
            ptx::umma_commit(&o_full[g * 2 + (cp & 1)]);
            if (g == 1) ptx::umma_commit(&kv_empty[cp & 1]);
        };
        for (int c = 0; c <= total; ++c) {
            const bool has_s = c < total;
            const int j = has_s ? c % nb : 0;
            if (has_s && j == 0 && c > 0) {
                // first block of the next item: the previous item's last P V products must not wait for the new Q tiles
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::mbar_wait(&p_ready[g], (c - 1) & 1);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) issue_pv(c - 1, g);
                    __syncwarp();
                }
                ptx::mbar_wait(q_full, (uint32_t)((c / nb) & 1));
                ptx::mbar_wait(&kv_full[c & 1], (c >> 1) & 1);
                ptx::tc_fence_after();
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (ptx::elect_one()) issue_s(c, g, nb == 1);
                    __syncwarp();
                }
                continue;
            }
            if (has_s) {
                if (c == 0) ptx::mbar_wait(q_full, 0);
                ptx::mbar_wait(&kv_full[c & 1], (c >> 1) & 1);
                ptx::tc_fence_after();
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (c > 0) {                                      // P_g(c-1) is in shared memory and S_g has been drained
                    ptx::mbar_wait(&p_ready[g], (c - 1) & 1);
                    ptx::tc_fence_after();
                }
                if (ptx::elect_one()) {
                    if (has_s) issue_s(c, g, j == nb - 1);
                    if (c > 0) issue_pv(c - 1, g);
                }
                __syncwarp();
            }
        }
    } else {
        const int g = (warp - 2) >> 2;                            // which query tile
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may touch
        const int r = quarter * 32 + lane;                        // query row of this thread within the tile
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t ts = tmem_S + (uint32_t)g * 128u + lane_off;
        const uint32_t to = tmem_O + (uint32_t)g * 128u + lane_off;
        uint64_t* my_s_full = &s_full[g];
        uint64_t* my_p_ready = &p_ready[g];
        uint64_t* my_o_full = &o_full[g * 2];
        uint8_t* myP = sP + g * 2 * FA_P_BYTES;
        const uint32_t p_row = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        float o[64];

        auto fold_o = [&](int c, float alpha) {                   // o = o * alpha + O_c (c = running block index)
            ptx::mbar_wait(&my_o_full[c & 1], (c >> 1) & 1);
            ptx::tc_fence_after();
            float t[64];
            ptx::tmem_ld32(to + (uint32_t)(c & 1) * 64u, t);
            ptx::tmem_ld32(to + (uint32_t)(c & 1) * 64u + 32u, t + 32);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = fmaf(o[i], alpha, t[i]);
        };

        int c = 0;                                                // running key-block index of this CTA
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int bh = item / n_qt, q0 = (item - bh * n_qt) * FA_BQ;
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = 0.f;
            float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
            for (int j = 0; j < nb; ++j, ++c) {
                ptx::mbar_wait(my_s_full, c & 1);
                ptx::tc_fence_after();
                const int kbase = j * FA_BK;
                const bool tail = kbase + FA_BK > T;
                float s[64];
                // ---- pass 1: row max over the 128 scores --------------------------------------------------------
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    ptx::tmem_ld32(ts + hf * 64, s);
                    ptx::tmem_ld32(ts + hf * 64 + 32, s + 32);
                    ptx::tmem_ld_wait();
                    if (tail) {
#pragma unroll
                        for (int cc = 0; cc < 64; ++cc) if (kbase + hf * 64 + cc >= T) s[cc] = -INFINITY;
                    }
#pragma unroll
                    for (int cc = 0; cc < 64; cc += 4) {
                        mx0 = fmaxf(mx0, s[cc]); mx1 = fmaxf(mx1, s[cc + 1]); mx2 = fmaxf(mx2, s[cc + 2]); mx3 = fmaxf(mx3, s[cc + 3]);
                    }
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                const float m_new = fmaxf(m, mx);
                const float alpha = ex2((m - m_new) * LOG2E);     // 0 on the first block (m = -inf)
                const float nm = -m_new * LOG2E;
                m = m_new;
                // ---- pass 2: p = exp(s - m), h16 P tile into swizzled shared memory -----------------------------------
                float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
                const uint32_t prow_s = ptx::smem_u32(myP + (c & 1) * FA_P_BYTES + p_row);
#pragma unroll
                for (int hf = 1; hf >= 0; --hf) {                  // second half first: it is already in registers
                    if (hf == 0) {
                        ptx::tmem_ld32(ts, s);
                        ptx::tmem_ld32(ts + 32, s + 32);
                        ptx::tmem_ld_wait();
                        if (tail) {
#pragma unroll
                            for (int cc = 0; cc < 64; ++cc) if (kbase + cc >= T) s[cc] = -INFINITY;
                        }
                    }
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {               // 8 keys = one 16-byte piece of the swizzled row
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) p[i] = ex2(fmaf(s[c8 * 8 + i], LOG2E, nm));
                        sum0 += p[0] + p[4]; sum1 += p[1] + p[5]; sum2 += p[2] + p[6]; sum3 += p[3] + p[7];
                        uint4 u;
                        u.x = pack_h16x2(p[0], p[1]); u.y = pack_h16x2(p[2], p[3]);
                        u.z = pack_h16x2(p[4], p[5]); u.w = pack_h16x2(p[6], p[7]);
                        ptx::sts128(prow_s + hf * FA_TILE_BYTES + (((uint32_t)c8 ^ sw) << 4), u);
                    }
                }
                l = fmaf(l, alpha, (sum0 + sum1) + (sum2 + sum3));
                ptx::fence_proxy_async();                         // P visible to the tensor core's async proxy
                ptx::tc_fence_before();                           // our TMEM reads of S are complete
                ptx::mbar_arrive(my_p_ready);
                if (j > 0) fold_o(c - 1, alpha_prev);
                alpha_prev = alpha;
            }
            fold_o(c - 1, alpha_prev);
            const int t = q0 + g * 128 + r;
            if (t < T) {
                const float inv = 1.0f / l;
                const int b = bh / H, h = bh - b * H;
                h16* dst = out + ((size_t)b * T + t) * ((size_t)H * 64) + (size_t)h * 64;
#pragma unroll
                for (int i = 0; i < 64; i += 8) {
                    uint4 u;
                    u.x = pack_h16x2(o[i] * inv, o[i + 1] * inv); u.y = pack_h16x2(o[i + 2] * inv, o[i + 3] * inv);
                    u.z = pack_h16x2(o[i + 4] * inv, o[i + 5] * inv); u.w = pack_h16x2(o[i + 6] * inv, o[i + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + i) = u;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

int make_map_3d(CUtensorMap* map, const void* base, int T, int BH) {
    cuuint64_t dims[3] = {64, (cuuint64_t)T, (cuuint64_t)BH};
    cuuint64_t strides[2] = {128, (cuuint64_t)T * 128};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_fa(map, WIPA_H16_TMA_TYPE, 3, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wipa_set_error("enc_attention_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return WIPA_ECUDA;
    }
    return WIPA_OK;
}

}  // namespace

// q, k, v: h16 [B, H, T, 64] (q pre-scaled); out: h16 [B, T, H*64]
int launch_enc_attention_tc(const h16* q, const h16* k, const h16* v, h16* out, int B, int H, int T, cudaStream_t st) {
    if (g_encode_fa == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WIPA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WIPA_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, WIPA_ECUDA, "cuTensorMapEncodeTiled not available");
        g_encode_fa = reinterpret_cast<EncodeTiledFn>(fn);
    }
    WIPA_CHECK(T >= 1 && B >= 1 && H >= 1, WIPA_EINVAL, "enc_attention_tc: bad shape");
    CUtensorMap tmQ, tmK, tmV;
    WIPA_TRY(make_map_3d(&tmQ, q, T, B * H));
    WIPA_TRY(make_map_3d(&tmK, k, T, B * H));
    WIPA_TRY(make_map_3d(&tmV, v, T, B * H));
    dim3 grid(cdiv(T, FA_BQ), B * H);
    static const int form = getenv("WIPA_FA_FORM") ? atoi(getenv("WIPA_FA_FORM")) : 2;
    if (form == 2) {
        static SmemAttr attr2;
        static int n_sm = 0;
        if (n_sm == 0) {
            int dev = 0;
            WIPA_CUDA_CHECK(cudaGetDevice(&dev));
            WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        }
        WIPA_TRY(wipa_ensure_smem(enc_attention_tc_persistent_kernel, (size_t)FA_SMEM, attr2));
        const int n_items = (int)(grid.x * grid.y);
        enc_attention_tc_persistent_kernel<<<n_items < n_sm ? n_items : n_sm, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, out, H, T, n_items);
        WIPA_LAUNCHED();
        return WIPA_OK;
    }
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(enc_attention_tc_kernel, (size_t)FA_SMEM, attr));
    enc_attention_tc_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, out, H, T);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
