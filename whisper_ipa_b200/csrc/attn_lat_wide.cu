// Latent cross-attention (see attn_lat.cu for the algebra and the kernel it is derived from) for the wide models: 16 heads
// (whisper-medium, d = 1024) and 20 heads (whisper-large / large-v2 / large-v3, d = 1280; ref:scripts/transcribe_single.py:12
// hard-codes large-v3).
//
// attn_lat.cu gives every 64-column tile of E its own warp and puts the heads on the 16 rows of the mma.sync tile.  At 16 heads
// that is 17 warps of 96 registers (spills) and a 2-deep ring: measured 6.8 cycles per mma.sync per SM against 4.85 here, and
// tensor-bound (206 us per launch at 64 utterances x 5 beams whether E comes from HBM or L2).  With 20 heads it would be two
// row tiles per warp and 21 warps, and [32 x 1280] fp32 accumulators do not fit the register file.  Here a warp owns TWO column
// tiles (128 columns), so 8 / 10 consumer warps with up to 168 registers, no spills, and room for a 3-deep ring at 16 heads.
// 20 heads are split into two groups of 10, one CTA per group:
//
//   * the two CTAs of a pair (blockIdx.x = 2 i + group) walk the SAME (sequence, chunk) ranges at the same pace, so every E
//     chunk is requested by both within a few hundred cycles: one of the two reads comes from HBM, the other from L2.  No
//     cluster, no multicast, no barrier between the two - HBM traffic stays ~1 x E, the L2 -> SM traffic doubles;
//   * a CTA has NW = 10 (8 at 16 heads) consumer warps; warp w owns column tiles 2 w and 2 w + 1 (128 columns) for both
//     products and does the softmax bookkeeping of heads h0 + w, h0 + w + NW, ..  Row tile = the group's heads (10 of 16 rows
//     used at 20 heads, all 16 at 16 heads);
//   * partials / counters of the stream-K split are per (sequence, group);
//   * beam search: like attn_lat.cu, the K beams of an utterance are walked by K CTAs (x G groups) in lockstep, so E leaves HBM
//     once per utterance (blockIdx.x = (range * K + beam) * G + group).
//
// Only the chunk-tiled, pre-swizzled E layout (common.cuh lat_tile_offset, 32 keys per chunk) is supported.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int XW_KEYS = 32;
constexpr int XW_PITCH = XW_KEYS + 8;         // floats per partial-score row / h16 per P row
constexpr int XW_BULK_SPLIT = 4;
constexpr float XW_LOG2E = 1.4426950408889634f;

// H heads = 64-column tiles of E; G head groups = CTAs per (range, beam); NW consumer warps; ST ring stages
template <int H_, int G_, int NW_, int ST_>
struct XwCfg {
    static constexpr int H = H_, G = G_, NW = NW_, ST = ST_;
    static constexpr int NH = H / G;              // heads per CTA (rows of the mma tile in use)
    static constexpr int CT = H / NW;             // column tiles per warp
    static constexpr int HPW = (NH + NW - 1) / NW;   // heads whose softmax a warp keeps
    static constexpr int D = H * 64;
    static constexpr uint32_t STAGE_BYTES = (uint32_t)XW_KEYS * H * 128u;
    static constexpr int ML_OFF = NW * 32 * CT * 32;          // [warp][lane][CT x 32 accumulators], then m[16] + l[16]
    static constexpr int SLOT_FLOATS = ML_OFF + 32;
    static constexpr size_t SMEM = (size_t)ST * STAGE_BYTES + (size_t)NW * NH * XW_PITCH * 4 + 16 * XW_PITCH * 2 + 32 * 4 + 64 + 1024;
    static_assert(H % G == 0 && H % NW == 0 && NH <= 16 && CT == 2, "shape");
};
using Xw20 = XwCfg<20, 2, 10, 2>;
using Xw16 = XwCfg<16, 1, 8, 3>;

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
template <int NTHR> __device__ __forceinline__ void xw_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory"); }
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
// warp maximum of floats in one redux.sync (IEEE floats order like sign-magnitude integers)
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

// normalise and store this thread's accumulators: rows g / g + 8 of the group (heads h0 + g, h0 + g + 8), the warp's 128 columns
template <class Cfg>
__device__ __forceinline__ void xw_store(h16* __restrict__ Cout, int s, int h0, int w, int g, int t, bool row_lo, bool row_hi,
                                         const float (&acc)[Cfg::CT][8][4], float il_lo, float il_hi) {
    h16* c_lo = Cout + ((size_t)s * Cfg::H + h0 + g) * Cfg::D + w * (Cfg::CT * 64) + 2 * t;
    h16* c_hi = c_lo + (size_t)8 * Cfg::D;
#pragma unroll
    for (int ct = 0; ct < Cfg::CT; ++ct) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (row_lo) *reinterpret_cast<uint32_t*>(c_lo + ct * 64 + j * 8) = pack_h16x2(acc[ct][j][0] * il_lo, acc[ct][j][1] * il_lo);
            if (row_hi) *reinterpret_cast<uint32_t*>(c_hi + ct * 64 + j * 8) = pack_h16x2(acc[ct][j][2] * il_hi, acc[ct][j][3] * il_hi);
        }
    }
}

template <class Cfg>
__global__ void __launch_bounds__((Cfg::NW + 1) * 32, 1)
cross_attention_latent_wide_kernel(const h16* __restrict__ Et, const h16* __restrict__ Qp, const int* __restrict__ utt_of_seq,
                                   h16* __restrict__ Cout, int S, int T, float* __restrict__ part, int* __restrict__ counters,
                                   int slots_per_seq, int K) {
    constexpr int H = Cfg::H, G = Cfg::G, NW = Cfg::NW, NH = Cfg::NH, CT = Cfg::CT, HPW = Cfg::HPW, D = Cfg::D, ST = Cfg::ST;
    constexpr uint32_t STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int NTHR = NW * 32;
    extern __shared__ uint8_t xw_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xw_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sE = smem;                                                      // [stage][H tiles][KEYS][128 B], 128B-swizzled
    float* Sp = reinterpret_cast<float*>(sE + ST * STAGE_BYTES);             // [warp][head of the group][PITCH]
    h16* Pm = reinterpret_cast<h16*>(Sp + NW * NH * XW_PITCH);               // [16][PITCH]
    float* alpha = reinterpret_cast<float*>(Pm + 16 * XW_PITCH);             // [16] rescale of C
    uint64_t* full = reinterpret_cast<uint64_t*>(alpha + 32);
    uint64_t* empty = full + ST;
    int* last_flag = reinterpret_cast<int*>(empty + ST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = (range * K + beam) * G + group: the K * G CTAs of a range walk the same (utterance slot, chunk) units
    const int grp = (int)blockIdx.x % G;
    const int bk = (int)blockIdx.x / G;
    const int bx = bk / K, beam = bk - bx * K, nx = (int)gridDim.x / (G * K);
    const int h0 = grp * NH;
    const int n_chunks = (T + XW_KEYS - 1) / XW_KEYS;
    // stream-K: the (utterance slot, chunk) units form one list cut into equal contiguous ranges
    const long long n_units = (long long)(S / K) * n_chunks;
    const long long u_lo = n_units * bx / nx, u_hi = n_units * (bx + 1) / nx;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], (uint32_t)NW); }
        ptx::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 16 * XW_PITCH; i += blockDim.x) Pm[i] = f32_to_h16(0.f);
    {
        // the ragged last chunk of a sequence copies only its valid keys; the rest of the stage then holds whatever an earlier
        // chunk left there (finite values, multiplied by p = 0) - but never uninitialised shared memory
        uint4* z = reinterpret_cast<uint4*>(sE);
        for (int i = threadIdx.x; i < (int)(ST * STAGE_BYTES / 16); i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    if (threadIdx.x < 32) alpha[threadIdx.x] = 1.f;
    __syncthreads();

    if (warp == NW) {
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
                    uint8_t* dst = sE + st * STAGE_BYTES;
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(Et) + ((size_t)u * n_chunks + ch) * STAGE_BYTES;
                    const int valid = (T - ch * XW_KEYS) < XW_KEYS ? (T - ch * XW_KEYS) : XW_KEYS;
                    if (valid == XW_KEYS) {                      // the whole chunk: contiguous, a few large bulk copies
                        constexpr uint32_t piece = STAGE_BYTES / XW_BULK_SPLIT;
                        ptx::mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
#pragma unroll
                        for (int i = 0; i < XW_BULK_SPLIT; ++i) ptx::bulk_load_1d(dst + i * piece, src + i * piece, piece, &full[st]);
                    } else {                                     // ragged end of the sequence: the valid rows of each column tile
                        ptx::mbar_arrive_expect_tx(&full[st], (uint32_t)(H * valid * 128));
                        for (int h = 0; h < H; ++h)
                            ptx::bulk_load_1d(dst + h * (XW_KEYS * 128), src + h * (XW_KEYS * 128), (uint32_t)(valid * 128), &full[st]);
                    }
                }
                __syncwarp();
                if (++st == ST) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers -----------------------------------------------------------------------------------------------------
    const int w = warp;
    const int g = lane >> 2, t = lane & 3;
    const bool row_lo = g < NH, row_hi = g + 8 < NH;
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
        // A fragments of Q': rows g / g + 8 (heads h0 + g, h0 + g + 8), this warp's 128 columns = 2 tiles x 4 k-steps of 16
        uint32_t qa[CT][4][4];
        {
            const uint32_t* q_lo = reinterpret_cast<const uint32_t*>(Qp + ((size_t)s * H + h0 + g) * D + w * (CT * 64) + 2 * t);
            const uint32_t* q_hi = q_lo + (size_t)8 * D / 2;
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    qa[ct][ks][0] = row_lo ? q_lo[ct * 32 + ks * 8] : 0u;          // 16 h16 = 8 words per k-step
                    qa[ct][ks][1] = row_hi ? q_hi[ct * 32 + ks * 8] : 0u;
                    qa[ct][ks][2] = row_lo ? q_lo[ct * 32 + ks * 8 + 4] : 0u;      // columns + 8
                    qa[ct][ks][3] = row_hi ? q_hi[ct * 32 + ks * 8 + 4] : 0u;
                }
            }
        }
        float acc[CT][8][4];
#pragma unroll
        for (int ct = 0; ct < CT; ++ct)
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[ct][j][0] = 0.f; acc[ct][j][1] = 0.f; acc[ct][j][2] = 0.f; acc[ct][j][3] = 0.f; }
        float m_run[HPW];                           // running maxima of heads h0 + w + i NW (replicated over the lanes of warp w)
#pragma unroll
        for (int i = 0; i < HPW; ++i) m_run[i] = -INFINITY;
        float accl[4] = {0.f, 0.f, 0.f, 0.f};      // l = sum_t p[t] as P x ones on the tensor pipe: rows g / g + 8 in [0] / [2]

        for (int ch = ch0; ch < ch1; ++ch) {
            ptx::mbar_wait(&full[st], ph);
            const uint32_t tile0 = sE_s + (uint32_t)st * STAGE_BYTES + (uint32_t)(w * CT) * (XW_KEYS * 128);
            // 1. partial scores over this warp's 128 columns
            {
                float sc[XW_KEYS / 8][4];
#pragma unroll
                for (int nt = 0; nt < XW_KEYS / 8; ++nt) { sc[nt][0] = 0.f; sc[nt][1] = 0.f; sc[nt][2] = 0.f; sc[nt][3] = 0.f; }
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    const uint32_t tile = tile0 + (uint32_t)ct * (XW_KEYS * 128);
                    uint32_t bfr[XW_KEYS / 8][2][4];
#pragma unroll
                    for (int nt = 0; nt < XW_KEYS / 8; ++nt) {
                        const int key = nt * 8 + (lane & 7);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int c16 = half * 4 + (lane >> 3);
                            ldmatrix_x4(tile + (uint32_t)key * 128u + (uint32_t)((c16 ^ (key & 7)) << 4), bfr[nt][half][0], bfr[nt][half][1],
                                        bfr[nt][half][2], bfr[nt][half][3]);
                        }
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                        for (int nt = 0; nt < XW_KEYS / 8; ++nt)
                            mma_h16(sc[nt], qa[ct][ks], bfr[nt][ks >> 1][(ks & 1) * 2], bfr[nt][ks >> 1][(ks & 1) * 2 + 1]);
                    }
                }
#pragma unroll
                for (int nt = 0; nt < XW_KEYS / 8; ++nt) {
                    if (row_lo) sts_f32x2(Sp_s + (uint32_t)((w * NH + g) * XW_PITCH + nt * 8 + 2 * t) * 4u, sc[nt][0], sc[nt][1]);
                    if (row_hi) sts_f32x2(Sp_s + (uint32_t)((w * NH + g + 8) * XW_PITCH + nt * 8 + 2 * t) * 4u, sc[nt][2], sc[nt][3]);
                }
            }
            xw_bar<NTHR>();
            // fragments of E for step 3 (first column tile) do not depend on the softmax: fetch them behind the reduction
            uint32_t vfr[XW_KEYS / 16][4][4];
            auto load_vfr = [&](uint32_t tile) {
#pragma unroll
                for (int ks = 0; ks < XW_KEYS / 16; ++ks) {
                    const int key = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        const int c16 = jp * 2 + (lane >> 4);
                        ldmatrix_x4_trans(tile + (uint32_t)key * 128u + (uint32_t)((c16 ^ (key & 7)) << 4), vfr[ks][jp][0], vfr[ks][jp][1],
                                          vfr[ks][jp][2], vfr[ks][jp][3]);
                    }
                }
            };
            load_vfr(tile0);
            // 2. heads h0 + w (+ NW ..): sum the partials of the NW warps, maximum by redux.sync, p in h16, rescale factor
#pragma unroll
            for (int i = 0; i < HPW; ++i) {
                const int hh = w + i * NW;                         // head of the group (warp-uniform)
                if (hh < NH) {
                    const uint32_t row = Sp_s + (uint32_t)(hh * XW_PITCH + lane) * 4u;
                    float v0a = 0.f, v0b = 0.f;
#pragma unroll
                    for (int ww = 0; ww < NW; ww += 2) {
                        v0a += lds_f32(row + (uint32_t)(ww * NH * XW_PITCH) * 4u);
                        v0b += lds_f32(row + (uint32_t)((ww + 1) * NH * XW_PITCH) * 4u);
                    }
                    float v0 = v0a + v0b;
                    if (ch * XW_KEYS + lane >= T) v0 = -INFINITY;
                    const float m_new = fmaxf(m_run[i], warp_max_f32(v0));             // finite: every chunk holds a valid key
                    const float mb = m_new * XW_LOG2E;
                    const float p0 = ex2_ftz(fmaf(v0, XW_LOG2E, -mb));
                    const float a = ex2_ftz(fmaf(m_run[i], XW_LOG2E, -mb));
                    m_run[i] = m_new;
                    sts_b16(Pm_s + (uint32_t)(hh * XW_PITCH + lane) * 2u, f32_to_h16(p0));
                    if (lane == 0) alpha[hh] = a;
                }
            }
            xw_bar<NTHR>();
            // 3. C = alpha * C + P E over this warp's columns
            {
                const float a_lo = alpha[g], a_hi = alpha[g + 8];
                accl[0] *= a_lo; accl[1] *= a_lo; accl[2] *= a_hi; accl[3] *= a_hi;
                uint32_t pa[XW_KEYS / 16][4];
#pragma unroll
                for (int ks = 0; ks < XW_KEYS / 16; ++ks) {
                    const uint32_t p_lo = Pm_s + (uint32_t)(g * XW_PITCH + ks * 16 + 2 * t) * 2u;
                    const uint32_t p_hi = p_lo + 8u * XW_PITCH * 2u;
                    pa[ks][0] = lds32(p_lo); pa[ks][1] = lds32(p_hi); pa[ks][2] = lds32(p_lo + 16u); pa[ks][3] = lds32(p_hi + 16u);
                }
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    if (ct > 0) load_vfr(tile0 + (uint32_t)ct * (XW_KEYS * 128));
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[ct][j][0] *= a_lo; acc[ct][j][1] *= a_lo; acc[ct][j][2] *= a_hi; acc[ct][j][3] *= a_hi; }
#pragma unroll
                    for (int ks = 0; ks < XW_KEYS / 16; ++ks) {
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) {
                            mma_h16(acc[ct][2 * jp], pa[ks], vfr[ks][jp][0], vfr[ks][jp][1]);
                            mma_h16(acc[ct][2 * jp + 1], pa[ks], vfr[ks][jp][2], vfr[ks][jp][3]);
                        }
                        if (ct == 0) mma_h16(accl, pa[ks], WIPA_H16_ONE_X2, WIPA_H16_ONE_X2);   // x ones: row sums of P
                    }
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[st]);
            if (++st == ST) { st = 0; ph ^= 1; }
        }

        // ---- end of the (sequence, chunk range) segment ---------------------------------------------------------------------
        if (ch0 == 0 && ch1 == n_chunks) {              // the whole sequence was ours: normalise and store
            xw_store<Cfg>(Cout, s, h0, w, g, t, row_lo, row_hi, acc, row_lo ? 1.f / accl[0] : 0.f, row_hi ? 1.f / accl[2] : 0.f);
            continue;
        }
        // a part of the sequence: leave (m, l, unnormalised C) in this range's slot of (sequence, group); whoever brings the chunk
        // count of (sequence, group) to n_chunks merges the slots in order (deterministic) and stores
        const long long first_unit = (long long)v * n_chunks;
        const int cta_first = (int)(((first_unit + 1) * nx + n_units - 1) / n_units) - 1;          // range holding chunk 0
        const int cta_last = (int)(((first_unit + n_chunks) * nx + n_units - 1) / n_units) - 1;   // range holding the last chunk
        float* seq_part = part + ((size_t)s * G + grp) * slots_per_seq * Cfg::SLOT_FLOATS;
        {
            float* slot = seq_part + (size_t)(bx - cta_first) * Cfg::SLOT_FLOATS;
            float4* dst = reinterpret_cast<float4*>(slot + (w * 32 + lane) * (CT * 32));
#pragma unroll
            for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[ct * 8 + j] = make_float4(acc[ct][j][0], acc[ct][j][1], acc[ct][j][2], acc[ct][j][3]);
            float* ml = slot + Cfg::ML_OFF;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < HPW; ++i) if (w + i * NW < NH) ml[w + i * NW] = m_run[i];
            }
            if (w == 0 && t == 0) {                                  // every warp holds the same l: warp 0 writes rows g / g + 8
                if (row_lo) ml[16 + g] = accl[0];
                if (row_hi) ml[16 + g + 8] = accl[2];
            }
        }
        __threadfence();
        xw_bar<NTHR>();
        if (threadIdx.x == 0) {
            const int mine = ch1 - ch0;
            const int old = atomicAdd(&counters[s * G + grp], mine);
            const bool last = old + mine == n_chunks;
            if (last) counters[s * G + grp] = 0;                 // ready for the next launch
            *last_flag = last ? 1 : 0;                           // next written after at least two more barriers
        }
        xw_bar<NTHR>();
        if (*last_flag == 0) continue;
        __threadfence();
        {
            const int n_slots = cta_last - cta_first + 1;
            constexpr int ml_off = Cfg::ML_OFF;
            float M_lo = -INFINITY, M_hi = -INFINITY;
            for (int i = 0; i < n_slots; ++i) {
                const float* sl = seq_part + (size_t)i * Cfg::SLOT_FLOATS + ml_off;
                if (row_lo) M_lo = fmaxf(M_lo, __ldcg(sl + g));
                if (row_hi) M_hi = fmaxf(M_hi, __ldcg(sl + g + 8));
            }
            float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
            for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[ct][j][0] = 0.f; acc[ct][j][1] = 0.f; acc[ct][j][2] = 0.f; acc[ct][j][3] = 0.f; }
            for (int i = 0; i < n_slots; ++i) {
                const float* sl = seq_part + (size_t)i * Cfg::SLOT_FLOATS;
                const float f_lo = row_lo ? ex2_ftz((__ldcg(sl + ml_off + g) - M_lo) * XW_LOG2E) : 0.f;
                const float f_hi = row_hi ? ex2_ftz((__ldcg(sl + ml_off + g + 8) - M_hi) * XW_LOG2E) : 0.f;
                if (row_lo) l_lo = fmaf(f_lo, __ldcg(sl + ml_off + 16 + g), l_lo);
                if (row_hi) l_hi = fmaf(f_hi, __ldcg(sl + ml_off + 16 + g + 8), l_hi);
                const float4* src = reinterpret_cast<const float4*>(sl + (w * 32 + lane) * (CT * 32));
#pragma unroll
                for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 vv = __ldcg(src + ct * 8 + j);
                        acc[ct][j][0] = fmaf(f_lo, vv.x, acc[ct][j][0]); acc[ct][j][1] = fmaf(f_lo, vv.y, acc[ct][j][1]);
                        acc[ct][j][2] = fmaf(f_hi, vv.z, acc[ct][j][2]); acc[ct][j][3] = fmaf(f_hi, vv.w, acc[ct][j][3]);
                    }
            }
            xw_store<Cfg>(Cout, s, h0, w, g, t, row_lo, row_hi, acc, row_lo ? 1.f / l_lo : 0.f, row_hi ? 1.f / l_hi : 0.f);
        }
    }
}

template <class Cfg>
int xw_launch(const h16* Qp, const h16* E, const int* utt_of_seq, h16* C, int S, int T, int n_sm, float* part, size_t part_floats,
              int* counters, int K, cudaStream_t st) {
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(cross_attention_latent_wide_kernel<Cfg>, Cfg::SMEM, attr));
    const int n_chunks = cdiv(T, XW_KEYS);
    const long long n_units = (long long)(S / K) * n_chunks;
    const int per = Cfg::G * K;                                  // CTAs per range
    const int ranges = n_units < n_sm / per ? (int)n_units : n_sm / per;
    // a sequence is cut by at most n_chunks / (shortest range) range boundaries
    const int slots_per_seq = n_chunks / (int)(n_units / ranges) + 2;
    WIPA_CHECK((size_t)Cfg::G * S * slots_per_seq * Cfg::SLOT_FLOATS <= part_floats, WIPA_EINVAL,
               "cross_attention_latent_wide: partial scratch too small for %d sequences", S);
    WIPA_CUDA_CHECK(wipa_launch_c(4, cross_attention_latent_wide_kernel<Cfg>, dim3(ranges * per), dim3((Cfg::NW + 1) * 32), Cfg::SMEM, st, E, Qp,
                                  utt_of_seq, C, S, T, part, counters, slots_per_seq, K));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

}  // namespace

// 20 heads always; 16 heads when WIPA_XL_WIDE16 is not 0 (attn_lat.cu keeps its own 16-head instantiation)
int cross_attention_latent_wide_supported(int H) {
    static const int wide16 = getenv("WIPA_XL_WIDE16") ? atoi(getenv("WIPA_XL_WIDE16")) : 1;
    return H == 20 || (H == 16 && wide16 != 0);
}

// floats of partial scratch that any launch with S <= max_seqs sequences can need on a device with n_sm SMs
size_t cross_attention_latent_wide_scratch_floats(int H, int max_seqs, int n_sm) {
    if (H == 16) return (size_t)Xw16::G * (size_t)(2 * n_sm + 3 * max_seqs + 64) * (size_t)Xw16::SLOT_FLOATS;
    return (size_t)Xw20::G * (size_t)(2 * n_sm + 3 * max_seqs + 64) * (size_t)Xw20::SLOT_FLOATS;
}

// Qp: h16 [S, H, 64 H] absorbed queries; E: encoder output in the chunk-tiled layout ([U][chunk][tile][key][64 swizzled], 32
// keys per chunk); utt_of_seq: int [S]; C: h16 [S, H, 64 H]; part: cross_attention_latent_wide_scratch_floats; counters:
// int [2 S] zeroed once (self-resetting); beams: 1, or K (2..8) when utt_of_seq[s] = s / K
int launch_cross_attention_latent_wide(const h16* Qp, const h16* E, const int* utt_of_seq, h16* C, int S, int H, int T, float* part,
                                       size_t part_floats, int* counters, cudaStream_t st, int beams) {
    WIPA_CHECK(H == 20 || H == 16, WIPA_EUNSUPPORTED, "cross_attention_latent_wide: %d heads (16 or 20)", H);
    WIPA_CHECK(S >= 1 && T >= 1 && part && counters, WIPA_EINVAL, "cross_attention_latent_wide: bad argument");
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WIPA_CUDA_CHECK(cudaGetDevice(&dev));
        WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int K = (beams >= 2 && beams <= 8 && S % beams == 0) ? beams : 1;
    if (H == 16) return xw_launch<Xw16>(Qp, E, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
    return xw_launch<Xw20>(Qp, E, utt_of_seq, C, S, T, n_sm, part, part_floats, counters, K, st);
}
