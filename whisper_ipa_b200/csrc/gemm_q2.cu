// Absorbed cross-attention queries in ONE node (16-bit decode path, latent cross-attention):
//
//     q_h  = LN(x) Wq_h^T + bq_h            [S, 64]   per head h   (K = d; the 2^-3 scaling lives in Wq / bq)
//     q'_h = q_h Wk_h                        [S, d]                (K = 64)
//
// ctx.cu used to run these as two GEMM nodes (gemm_tc.cu: the d x d projection, then the head-batched K = 64 step) with q
// making a round trip through L2.  Each node of the decode step costs ~6 us of dependent latency whatever its size
// (DESIGN.md 3.2), so here the q tile never leaves the SM:
//
//   CTA (column chunk j, M tile, head h), 192 threads:
//     warp 0     TMA producer: the W2 chunk WkT_h[j] ([NC, 64], static: before griddepcontrol.wait), then the phase-1 ring of
//                A (128 x 64) and Wq_h (64 x 64) k-blocks, 128B-swizzled
//     warp 1     tcgen05 issuer: phase 1  acc1[128 x 64]  = A Wq_h^T        (TMEM columns [0, 64))
//                                phase 2  acc2[128 x NC]  = q_h WkT_h[j]^T  (TMEM columns [64, 64 + NC)), A = the h16 q tile the
//                                epilogue warps left in shared memory in the K-major 128B-swizzled layout
//     warps 2-5  acc1 -> (folded LayerNorm finish) + bias -> h16 -> swizzled shared tile; then acc2 -> h16 -> Q'[s, h, chunk]
//
// Every chunk CTA of a (M tile, head) recomputes the same q tile (d / NC times the phase-1 loads, all from L2): that buys
// d / NC times the CTAs, i.e. one full wave of 144 CTAs for whisper-small at 256 and 512 sequences.
#define WIPA_PDL_CLASS 8
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int Q2_BM = 128;
constexpr int Q2_BK = 64;
constexpr int Q2_STAGES = 3;
constexpr int Q2_KPB = 2;                                    // k-blocks per ring slot (one barrier round trip per 128 of K)
constexpr int Q2_A_BYTES = Q2_BM * Q2_BK * 2;                // 16 KB
constexpr int Q2_W_BYTES = 64 * Q2_BK * 2;                   // 8 KB: the 64 rows of Wq that belong to one head
constexpr int Q2_STAGE_A = Q2_KPB * Q2_A_BYTES;
constexpr int Q2_STAGE_W = Q2_KPB * Q2_W_BYTES;

// a CTA's chunk of NC output columns = NSUB sub-chunks of NW columns (one tcgen05.mma shape of N = NW <= 256 each)
template <int NW, int NSUB> struct Q2Cfg {
    static constexpr int NC = NW * NSUB;
    static constexpr int W2_BYTES = NC * Q2_BK * 2;          // WkT_h chunk [NC, 64]
    static constexpr int TMEM_COLS = 64 + NC <= 256 ? 256 : 512;  // 64 + NC columns, rounded up to a power of two
    static_assert(64 + NC <= 512 && NW <= 256 && NW % 16 == 0, "TMEM / UMMA shape");
    static constexpr int SMEM = Q2_STAGES * (Q2_STAGE_A + Q2_STAGE_W) + Q2_A_BYTES /*q tile*/ + W2_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn q2_encode = nullptr;

struct Q2Params {
    const float* bias;       // [d]: bq (beta-folded when ln_stats is given), indexed 64 h + j
    const float* ln_stats;   // folded LayerNorm (common.cuh): per-row (mean, M2) pieces of the residual, or nullptr
    const float* ln_c;       // [d] column sums of the gain-folded Wq
    int ln_nt;
    h16* out;                // Q' [S, H, d]
    int S, H, d;
};

template <int NW, int NSUB>
__global__ void __launch_bounds__(192)
xlq_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWq,
                 const __grid_constant__ CUtensorMap tmWkT, int num_kb, Q2Params p) {
    using Cfg = Q2Cfg<NW, NSUB>;
    constexpr int NC = Cfg::NC;
    extern __shared__ uint8_t q2_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q2_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = sA + Q2_STAGES * Q2_STAGE_A;
    uint8_t* sQ = sW + Q2_STAGES * Q2_STAGE_W;                 // q tile [128][64] h16, K-major, 128B-swizzled
    uint8_t* sW2 = sQ + Q2_A_BYTES;                            // WkT_h chunk [NC][64]
    uint64_t* full = reinterpret_cast<uint64_t*>(sW2 + Cfg::W2_BYTES);
    uint64_t* empty = full + Q2_STAGES;
    uint64_t* w2_full = empty + Q2_STAGES;
    uint64_t* acc1_full = w2_full + 1;
    uint64_t* q_ready = acc1_full + 1;
    uint64_t* acc2_full = q_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_full + 1);
    __shared__ float s_bias[64];
    __shared__ float s_lnc[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * NC;                            // first output column of this CTA's chunk
    const int t0 = blockIdx.y * Q2_BM;                         // first sequence of the M tile
    const int h = blockIdx.z;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmWq);
        ptx::prefetch_tensormap(&tmWkT);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Q2_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
            ptx::mbar_init(w2_full, 1);
            ptx::mbar_init(acc1_full, 1);
            ptx::mbar_init(q_ready, 128);
            ptx::mbar_init(acc2_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // static operands first (weights do not depend on the previous kernel): the W2 chunk and the Wq tiles of the first slots
        const int num_g = (num_kb + Q2_KPB - 1) / Q2_KPB;
        const int pre = num_g < Q2_STAGES ? num_g : Q2_STAGES;
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(w2_full, (uint32_t)Cfg::W2_BYTES);
            for (int j = 0; j < NSUB; ++j) ptx::tma_load_2d(sW2 + j * (NW * Q2_BK * 2), &tmWkT, w2_full, 0, h * p.d + n0 + j * NW);
            for (int g = 0; g < pre; ++g) {
                const int nk = (num_kb - g * Q2_KPB) < Q2_KPB ? (num_kb - g * Q2_KPB) : Q2_KPB;
                ptx::mbar_arrive_expect_tx(&full[g], (uint32_t)nk * (Q2_A_BYTES + Q2_W_BYTES));
                for (int i = 0; i < nk; ++i)
                    ptx::tma_load_2d(sW + g * Q2_STAGE_W + i * Q2_W_BYTES, &tmWq, &full[g], (g * Q2_KPB + i) * Q2_BK, h * 64);
            }
        }
        __syncwarp();
        pdl_wait();
        pdl_launch_dependents();
        if (ptx::elect_one()) {
            for (int g = 0; g < pre; ++g) {
                const int nk = (num_kb - g * Q2_KPB) < Q2_KPB ? (num_kb - g * Q2_KPB) : Q2_KPB;
                for (int i = 0; i < nk; ++i)
                    ptx::tma_load_3d(sA + g * Q2_STAGE_A + i * Q2_A_BYTES, &tmA, &full[g], (g * Q2_KPB + i) * Q2_BK, t0, 0);
            }
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        for (int g = pre; g < num_g; ++g) {
            ptx::mbar_wait(&empty[s], ph);
            if (ptx::elect_one()) {
                const int nk = (num_kb - g * Q2_KPB) < Q2_KPB ? (num_kb - g * Q2_KPB) : Q2_KPB;
                ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)nk * (Q2_A_BYTES + Q2_W_BYTES));
                for (int i = 0; i < nk; ++i) {
                    ptx::tma_load_3d(sA + s * Q2_STAGE_A + i * Q2_A_BYTES, &tmA, &full[s], (g * Q2_KPB + i) * Q2_BK, t0, 0);
                    ptx::tma_load_2d(sW + s * Q2_STAGE_W + i * Q2_W_BYTES, &tmWq, &full[s], (g * Q2_KPB + i) * Q2_BK, h * 64);
                }
            }
            __syncwarp();
            if (++s == Q2_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ---- phase 1: acc1 = A Wq_h^T ---------------------------------------------------------------------------------
        constexpr uint32_t idesc1 = ptx::idesc_h16_f32(Q2_BM, 64);
        constexpr uint32_t idesc2 = ptx::idesc_h16_f32(Q2_BM, NW);
        const int num_g = (num_kb + Q2_KPB - 1) / Q2_KPB;
        const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA));
        const uint32_t w_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sW));
        int s = 0;
        uint32_t ph = 0;
        for (int g = 0; g < num_g; ++g) {
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const int nk = (num_kb - g * Q2_KPB) < Q2_KPB ? (num_kb - g * Q2_KPB) : Q2_KPB;
                const uint32_t a_lo = a_lo0 + (uint32_t)s * (Q2_STAGE_A >> 4);
                const uint32_t w_lo = w_lo0 + (uint32_t)s * (Q2_STAGE_W >> 4);
#pragma unroll
                for (int i = 0; i < Q2_KPB; ++i) {
                    if (i < nk) {
#pragma unroll
                        for (int k = 0; k < Q2_BK / 16; ++k)
                            ptx::umma_h16(tmem_base, ptx::smem_desc_sw128(a_lo + (uint32_t)i * (Q2_A_BYTES >> 4) + 2 * k),
                                          ptx::smem_desc_sw128(w_lo + (uint32_t)i * (Q2_W_BYTES >> 4) + 2 * k), idesc1,
                                          (g | i | k) != 0 ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == Q2_STAGES) { s = 0; ph ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(acc1_full);
        __syncwarp();
        // ---- phase 2: acc2 = q_h WkT_h[chunk]^T (K = 64: four K16 steps) -------------------------------------------------
        ptx::mbar_wait(w2_full, 0);
        ptx::mbar_wait(q_ready, 0);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
            const uint32_t q_lo = ptx::smem_desc_lo(ptx::smem_u32(sQ));
            const uint32_t w2_lo = ptx::smem_desc_lo(ptx::smem_u32(sW2));
#pragma unroll
            for (int j = 0; j < NSUB; ++j)
#pragma unroll
                for (int k = 0; k < Q2_BK / 16; ++k)
                    ptx::umma_h16(tmem_base + 64u + (uint32_t)(j * NW), ptx::smem_desc_sw128(q_lo + 2 * k),
                                  ptx::smem_desc_sw128(w2_lo + (uint32_t)j * ((NW * Q2_BK * 2) >> 4) + 2 * k), idesc2, k != 0 ? 1u : 0u);
            ptx::umma_commit(acc2_full);
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: TMEM lane quarter fixed by warp id % 4; thread = row ------------------------------------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                       // row of the tile
        const int srow = t0 + r;                                 // sequence
        const bool row_ok = srow < p.S;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        pdl_wait();                                              // statistics come from the kernel that produced the residual
        const bool ln = p.ln_stats != nullptr;
        if (threadIdx.x - 64 < 64) {
            const int j = threadIdx.x - 64;
            s_bias[j] = p.bias != nullptr ? p.bias[h * 64 + j] : 0.f;
            s_lnc[j] = ln ? p.ln_c[h * 64 + j] : 0.f;
        }
        float2 mr = make_float2(0.f, 1.f);
        if (ln && row_ok) mr = ln_row_stats(p.ln_stats + (long long)srow * p.ln_nt * 2, p.ln_nt);
        asm volatile("bar.sync 1, 128;" ::: "memory");           // s_bias / s_lnc visible to the four epilogue warps
        ptx::mbar_wait(acc1_full, 0);
        ptx::tc_fence_after();
        {
            float v[64];
            ptx::tmem_ld32(taddr, v);
            ptx::tmem_ld32(taddr + 32, v + 32);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float x = ln ? mr.y * fmaf(-mr.x, s_lnc[i], v[i]) : v[i];
                v[i] = x + s_bias[i];
            }
            // q row -> h16 -> the K-major 128B-swizzled tile the tensor core reads as its A operand: 8-row atoms of 1024 bytes,
            // 128 bytes per row, 16-byte piece c of row r at piece (c ^ (r & 7))
            const uint32_t row_s = ptx::smem_u32(sQ) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
            const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint4 u;
                u.x = pack_h16x2(v[c8 * 8], v[c8 * 8 + 1]); u.y = pack_h16x2(v[c8 * 8 + 2], v[c8 * 8 + 3]);
                u.z = pack_h16x2(v[c8 * 8 + 4], v[c8 * 8 + 5]); u.w = pack_h16x2(v[c8 * 8 + 6], v[c8 * 8 + 7]);
                ptx::sts128(row_s + (((uint32_t)c8 ^ sw) << 4), u);
            }
        }
        ptx::fence_proxy_async();                                // the q tile is visible to the tensor core's async proxy
        ptx::tc_fence_before();                                  // our TMEM reads of acc1 are complete
        ptx::mbar_arrive(q_ready);
        ptx::mbar_wait(acc2_full, 0);
        ptx::tc_fence_after();
        h16* dst = p.out + ((size_t)srow * p.H + h) * p.d + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < NC; c0 += 32) {
            float v[32];
            ptx::tmem_ld32(taddr + 64u + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 u;
                    u.x = pack_h16x2(v[i], v[i + 1]); u.y = pack_h16x2(v[i + 2], v[i + 3]);
                    u.z = pack_h16x2(v[i + 4], v[i + 5]); u.w = pack_h16x2(v[i + 6], v[i + 7]);
                    *reinterpret_cast<uint4*>(dst + c0 + i) = u;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

int q2_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = q2_encode(map, WIPA_H16_TMA_TYPE, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wipa_set_error("xlq_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return WIPA_ECUDA;
    }
    return WIPA_OK;
}

template <int NW, int NSUB>
int q2_launch(const CUtensorMap& tmA, const CUtensorMap& tmWq, const CUtensorMap& tmWkT, int num_kb, const Q2Params& p, cudaStream_t st) {
    using Cfg = Q2Cfg<NW, NSUB>;
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(xlq_fused_kernel<NW, NSUB>, (size_t)Cfg::SMEM, attr));
    const dim3 grid(p.d / Cfg::NC, cdiv(p.S, Q2_BM), p.H);
    WIPA_CUDA_CHECK(wipa_launch(xlq_fused_kernel<NW, NSUB>, grid, dim3(192), (size_t)Cfg::SMEM, st, tmA, tmWq, tmWkT, num_kb, p));
    WIPA_LAUNCHED();
    return WIPA_OK;
}

}  // namespace

// A: h16 [S, d] (LayerNorm output, or the raw residual in h16 when ln_stats is given); Wq: h16 [d, d] (gain-folded in the
// latter case); WkT: h16 [H, d, 64] (ctx.cu xlat_wkt_kernel); bias: [d]; out: h16 [S, H, d]
int launch_xlq_fused(const h16* A, const h16* Wq, const h16* WkT, const float* bias, const float* ln_stats, const float* ln_c,
                     int ln_nt, h16* out, int S, int H, cudaStream_t st) {
    const int d = 64 * H;
    WIPA_CHECK(A && Wq && WkT && out && S >= 1 && d % 128 == 0, WIPA_EINVAL, "xlq_fused: bad argument");
    if (q2_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WIPA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WIPA_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, WIPA_ECUDA, "cuTensorMapEncodeTiled not available");
        q2_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    // 256-column chunks once S needs more than two M tiles (keeps the grid at one wave), 128 otherwise
    // (measured, whisper-small, us / decode step with 128 | 256 | 384-column chunks: 2054 | 2049 | 2092 at 256 sequences,
    // 3677 | 3633 | 3660 at 512)
    const int nc = (S > 256 && d % 256 == 0) ? 256 : 128;
    CUtensorMap tmA, tmWq, tmWkT;
    {
        cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)S, 1};
        cuuint64_t strides[2] = {(cuuint64_t)d * 2, (cuuint64_t)S * d * 2};
        cuuint32_t box[3] = {Q2_BK, Q2_BM, 1};
        WIPA_TRY(q2_map(&tmA, A, 3, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)d};
        cuuint64_t strides[1] = {(cuuint64_t)d * 2};
        cuuint32_t box[2] = {Q2_BK, 64};
        WIPA_TRY(q2_map(&tmWq, Wq, 2, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)H * d};
        cuuint64_t strides[1] = {64 * 2};
        cuuint32_t box[2] = {Q2_BK, (cuuint32_t)nc};
        WIPA_TRY(q2_map(&tmWkT, WkT, 2, dims, strides, box));
    }
    Q2Params p;
    p.bias = bias; p.ln_stats = ln_stats; p.ln_c = ln_c; p.ln_nt = ln_nt; p.out = out; p.S = S; p.H = H; p.d = d;
    const int num_kb = d / Q2_BK;
    if (nc == 256) return q2_launch<256, 1>(tmA, tmWq, tmWkT, num_kb, p, st);
    return q2_launch<128, 1>(tmA, tmWq, tmWkT, num_kb, p, st);
}
