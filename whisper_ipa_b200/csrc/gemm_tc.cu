// h16 path GEMM on the 5th-generation tensor cores:  C[M,N] = A[M,K] * W[N,K]^T, fp32 accumulate in TMEM.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64), 128B-swizzled,
//               into a STAGES-deep shared-memory ring guarded by full/empty mbarriers (KPB k-blocks per ring slot)
//   warp 1      allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per k-block and
//               tcgen05.commit's the stage back to the producer; the last commit signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld the 128 x BN fp32 accumulator (one row per thread) and apply the shared
//               epilogue (bias / GELU / residual / head split / paged-KV scatter / fused vocab argmax)
//
// The A operand is described by a 3-D tensor map (K, rows-per-batch, batches) whose row stride may be smaller than
// K: a k=3 conv1d over a channels-last, zero-row-padded signal is then exactly this GEMM (no im2col buffer),
// and out-of-range rows / K tails are zero-filled by TMA.
#include <stdlib.h>

#define WIPA_PDL_CLASS 8
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;

// BOXM = rows of A fetched per k-block: 128, or 64 when the whole problem has <= 64 rows (decode steps at batch <= 64).
// With BOXM = 64 the MMA still reads a 128-row operand: rows 64..127 alias the next ring slot, produce garbage
// accumulator lanes 64..127, and are never loaded by the epilogue (row_ok), so the ring can be twice as deep.
template <int BN, int BOXM> struct TcCfg {
    static constexpr int A_BYTES = BOXM * TC_BK * 2;
    static constexpr int W_BYTES = BN * TC_BK * 2;
    // A ring slot holds KPB consecutive k-blocks behind ONE full/empty barrier pair.  The narrow tiles are the
    // latency-bound decode GEMMs: their MMAs are short (32 cycles each, bound by the A read from shared memory), so a
    // barrier round trip per 64-wide k-block (~150 cycles of try_wait + commit) would dominate the issue loop.
    static constexpr int KPB = BN == 32 ? (BOXM == 64 ? 4 : 2) : (BN == 64 ? 2 : 1);
#ifndef WIPA_TC_STAGES_32
#define WIPA_TC_STAGES_32 4
#endif
#ifndef WIPA_TC_STAGES_64
#define WIPA_TC_STAGES_64 3
#endif
    static constexpr int STAGES = BN == 32 ? (BOXM == 64 ? 3 : WIPA_TC_STAGES_32) : (BN == 64 ? WIPA_TC_STAGES_64 : (BN == 128 ? 3 : 4));
    static constexpr int STAGE_A = KPB * A_BYTES;
    static constexpr int STAGE_W = KPB * W_BYTES;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_W) + (TC_BM - BOXM) * TC_BK * 2 /*over-read*/ + 1024 /*align*/ + 512 /*barriers*/;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

#ifdef WIPA_GEMM_DBG
__device__ unsigned long long g_gemm_dbg[4096 * 8];
#define DBG_STAMP(i) do { g_gemm_dbg[(blockIdx.y * gridDim.x + blockIdx.x) % 4096 * 8 + (i)] = clock64(); } while (0)
#else
#define DBG_STAMP(i) do { } while (0)
#endif

// CL > 1: CL consecutive N tiles of one M tile form a thread-block cluster.  Every CTA fetches 1/CL of the rows of each A
// k-block and TMA-multicasts it to the whole cluster, so the activations cross L2 -> SM once per cluster instead of
// once per N tile (the decode GEMMs are A-traffic bound: 24-96 N tiles re-read the same 128 x K rows).  A ring slot may
// only be refilled when ALL CTAs of the cluster have consumed it: tcgen05.commit multicasts the slot release.
// MODE: the epilogue (EpiMode) fixed at compile time, or -1 for the run-time switch.  The decode step's nodes each get their
// own instantiation, so that an 8 us kernel does not page in the SASS of seven epilogues it never runs (the all-in-one
// instantiations were 11.6 k instructions; the specialised ones are 1 - 3 k).
template <int BN, int BOXM, int CL, int MODE>
__global__ void __launch_bounds__(192)
gemm_h16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int num_kb,
                    int tiles_per_batch, int a_rpb, EpiParams ep) {
    using Cfg = TcCfg<BN, BOXM>;
    const int mode = MODE >= 0 ? MODE : ep.mode;
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + Cfg::STAGES * Cfg::STAGE_A + (TC_BM - BOXM) * TC_BK * 2;
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + Cfg::STAGES * Cfg::STAGE_W);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tmem_full = empty + Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int batch = blockIdx.y / tiles_per_batch;
    const int wn0 = n0 + batch * ep.w_brows;                  // first W row of this tile (batched W: a block of rows per batch)
    const int t0 = (blockIdx.y - batch * tiles_per_batch) * TC_BM;
    // split-K: CTA z of gridDim.z takes k-blocks [kb0, kb0 + num_kb) of the num_kb_total
    const int num_kb_total = num_kb;
    const int kb_per = (num_kb_total + (int)gridDim.z - 1) / (int)gridDim.z;
    const int kb0 = (int)blockIdx.z * kb_per;
    num_kb = min(kb_per, num_kb_total - kb0);

    if (threadIdx.x == 0) DBG_STAMP(0);
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], CL); }
            ptx::mbar_init(tmem_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CL > 1) ptx::cluster_sync_all();                      // every CTA's barriers exist before any peer multicasts into them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
    constexpr uint16_t mc_mask = (uint16_t)((1u << CL) - 1u);
    constexpr int SLICE_ROWS = BOXM / CL;                     // rows of each A k-block this CTA fetches for the cluster
    constexpr int SLICE_BYTES = SLICE_ROWS * TC_BK * 2;
    auto load_a = [&](uint8_t* dst, uint64_t* bar, int kcoord) {
        if (CL > 1) ptx::tma_load_3d_mc(dst + cta_rank * SLICE_BYTES, &tmA, bar, kcoord, t0 + (int)cta_rank * SLICE_ROWS, batch, mc_mask);
        else ptx::tma_load_3d(dst, &tmA, bar, kcoord, t0, batch);
    };

    // Role warps run warp-uniformly and elect one lane per issue: ptxas then keeps descriptors and barrier addresses
    // in uniform registers and emits bare UTMALDG / UTCHMMA (under `if (lane == 0)` every such instruction is wrapped
    // in an ELECT / BRA.U.ANY loop plus R2UR moves, ~135 cycles per MMA issue as measured with clock stamps).
    if (warp == 0) {
        // weights are static: fill the ring with W tiles BEFORE waiting for the previous kernel, then the A tiles
        constexpr int KPB = Cfg::KPB;
        const int num_g = (num_kb + KPB - 1) / KPB;               // ring slots to fill in total
        const int pre = num_g < Cfg::STAGES ? num_g : Cfg::STAGES;
        if (ptx::elect_one()) {
            for (int g = 0; g < pre; ++g) {
                const int nk = (num_kb - g * KPB) < KPB ? (num_kb - g * KPB) : KPB;
                ptx::mbar_arrive_expect_tx(&full[g], (uint32_t)nk * (Cfg::A_BYTES + Cfg::W_BYTES));
                for (int i = 0; i < nk; ++i)
                    ptx::tma_load_2d(sW + g * Cfg::STAGE_W + i * Cfg::W_BYTES, &tmW, &full[g], (kb0 + g * KPB + i) * TC_BK, wn0);
            }
            DBG_STAMP(1);
        }
        __syncwarp();
        pdl_wait();
        pdl_launch_dependents();                                  // the next kernel may start its prologue now (trigger only
                                                                  // after our own dependency is met: at most two grids overlap)
        if (ptx::elect_one()) {
            DBG_STAMP(2);
            for (int g = 0; g < pre; ++g) {
                const int nk = (num_kb - g * KPB) < KPB ? (num_kb - g * KPB) : KPB;
                for (int i = 0; i < nk; ++i)
                    load_a(sA + g * Cfg::STAGE_A + i * Cfg::A_BYTES, &full[g], (kb0 + g * KPB + i) * TC_BK);
            }
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;                                       // parity of the ring pass that filled slot s last
        for (int g = pre; g < num_g; ++g) {
            ptx::mbar_wait(&empty[s], ph);
            if (ptx::elect_one()) {
                const int nk = (num_kb - g * KPB) < KPB ? (num_kb - g * KPB) : KPB;
                ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)nk * (Cfg::A_BYTES + Cfg::W_BYTES));
                for (int i = 0; i < nk; ++i) {
                    load_a(sA + s * Cfg::STAGE_A + i * Cfg::A_BYTES, &full[s], (kb0 + g * KPB + i) * TC_BK);
                    ptx::tma_load_2d(sW + s * Cfg::STAGE_W + i * Cfg::W_BYTES, &tmW, &full[s], (kb0 + g * KPB + i) * TC_BK, wn0);
                }
            }
            __syncwarp();
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = ptx::idesc_h16_f32(TC_BM, BN);
        constexpr int KPB = Cfg::KPB;
        const int num_g = (num_kb + KPB - 1) / KPB;
        const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA));
        const uint32_t w_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sW));
        int s = 0;
        uint32_t ph = 0;
        for (int g = 0; g < num_g; ++g) {
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                if (g == 0) DBG_STAMP(3);
                const int nk = (num_kb - g * KPB) < KPB ? (num_kb - g * KPB) : KPB;
                const uint32_t a_lo = a_lo0 + (uint32_t)s * (Cfg::STAGE_A >> 4);
                const uint32_t w_lo = w_lo0 + (uint32_t)s * (Cfg::STAGE_W >> 4);
#pragma unroll
                for (int i = 0; i < KPB; ++i) {
                    if (i < nk) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            ptx::umma_h16(tmem_base, ptx::smem_desc_sw128(a_lo + (uint32_t)i * (Cfg::A_BYTES >> 4) + 2 * k),
                                           ptx::smem_desc_sw128(w_lo + (uint32_t)i * (Cfg::W_BYTES >> 4) + 2 * k), idesc,
                                           (g | i | k) != 0 ? 1u : 0u);
                    }
                }
                // a slot is free once EVERY CTA of the cluster has consumed it; when the whole K range fits the ring nobody
                // ever waits for a slot, the release is not sent and the closing cluster barrier is not needed either
                if (CL > 1 && num_g > Cfg::STAGES) ptx::umma_commit_mc(&empty[s], mc_mask);
                else if (CL == 1) ptx::umma_commit(&empty[s]);      // frees the ring slot when these MMAs retire
            }
            __syncwarp();
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        if (ptx::elect_one()) {
            DBG_STAMP(4);
            ptx::umma_commit(tmem_full);                     // accumulator complete
        }
        __syncwarp();
    } else {
        // epilogue warps: TMEM lane quarter is fixed by warp id % 4
        const int quarter = warp & 3;
        const int t = t0 + quarter * 32 + lane;
        const bool row_ok = t < ep.M_rows;
        const int m = batch * a_rpb + t;
        pdl_wait();                                          // the epilogue reads / writes activations of earlier kernels
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if ((BN == 32 && mode != EPI_ARGMAX) || (BN == 64 && mode == EPI_RESADD)) {
            // Latency-bound decode tiles: fetch bias and residual for the whole BN-column row segment BEFORE the
            // accumulator is ready, so that after the last MMA only tcgen05.ld + adds + stores remain.
            __shared__ int s_last;
            const bool active = row_ok && !(BOXM == 64 && quarter >= 2);
            const bool fast = ep.vec_ok && (n0 + BN <= ep.N);
            float add[BN];
#pragma unroll
            for (int i = 0; i < BN; ++i) add[i] = 0.f;
            long long row = 0;
            // folded LayerNorm, consumer side: this row's (mean, rstd) from the producer's per-piece statistics and the
            // column sums of the gain-folded weights, all fetched before the accumulator is ready
            const bool ln = ep.ln_stats != nullptr;
            float2 mr = make_float2(0.f, 1.f);
            float lnc[BN];
#pragma unroll
            for (int i = 0; i < BN; ++i) lnc[i] = 0.f;
            if (ln && active) {
                mr = ln_row_stats(ep.ln_stats + (long long)m * ep.ln_nt * 2, ep.ln_nt);
                if (fast) {
#pragma unroll
                    for (int i = 0; i < BN; i += 4) {
                        const float4 c4 = *reinterpret_cast<const float4*>(ep.ln_c + n0 + i);
                        lnc[i] = c4.x; lnc[i + 1] = c4.y; lnc[i + 2] = c4.z; lnc[i + 3] = c4.w;
                    }
                }
            }
            if (active && fast) {
                if (ep.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < BN; i += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(ep.bias + wn0 + i);      // bias is indexed like the rows of W
                        add[i] = b.x; add[i + 1] = b.y; add[i + 2] = b.z; add[i + 3] = b.w;
                    }
                }
                if (mode == EPI_RESADD) {
                    const int ob = m / ep.o_rpb;
                    row = (long long)ob * ep.o_bstride + (long long)(m - ob * ep.o_rpb) * ep.ldo + n0;
#pragma unroll
                    for (int i = 0; i < BN; i += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(ep.resid + row + i);
                        add[i] += x.x; add[i + 1] += x.y; add[i + 2] += x.z; add[i + 3] += x.w;
                    }
                }
            }
            ptx::mbar_wait(tmem_full, 0);
            if (threadIdx.x == 64) DBG_STAMP(5);
            ptx::tc_fence_after();
            float v[BN];
            if (!(BOXM == 64 && quarter >= 2)) {                 // warp-uniform: tcgen05.ld is a warp-collective
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) ptx::tmem_ld32(taddr + c0, v + c0);
                ptx::tmem_ld_wait();
            }
            bool finish = true;
            if (gridDim.z > 1) {
                // ---- deterministic split-K: publish the raw partial, take a ticket, the last CTA sums in split order ----
                const int tile = blockIdx.y * gridDim.x + blockIdx.x;
                const int trow = quarter * 32 + lane;                       // row of the tile held by this thread
                float* mine = ep.sk_part + (((size_t)tile * gridDim.z + blockIdx.z) * TC_BM + trow) * BN;
                if (!(BOXM == 64 && quarter >= 2)) {
#pragma unroll
                    for (int i = 0; i < BN; i += 4) *reinterpret_cast<float4*>(mine + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");                // the four epilogue warps
                if (threadIdx.x == 64) {
                    const int t = atomicAdd(&ep.sk_count[tile], 1);
                    s_last = (t == (int)gridDim.z - 1);
                    if (s_last) ep.sk_count[tile] = 0;                        // ready for the next launch
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                finish = s_last != 0;
                if (finish && !(BOXM == 64 && quarter >= 2)) {
                    __threadfence();
#pragma unroll
                    for (int i = 0; i < BN; ++i) v[i] = 0.f;
                    for (int z = 0; z < (int)gridDim.z; ++z) {
                        const float* p = ep.sk_part + (((size_t)tile * gridDim.z + z) * TC_BM + trow) * BN;
#pragma unroll
                        for (int i = 0; i < BN; i += 4) {
                            const float4 x = __ldcg(reinterpret_cast<const float4*>(p + i));
                            v[i] += x.x; v[i + 1] += x.y; v[i + 2] += x.z; v[i + 3] += x.w;
                        }
                    }
                }
            }
            if (finish && !(BOXM == 64 && quarter >= 2)) {
                if (!active) {
                    // padding row of the M tile: nothing to store
                } else if (fast) {
                    if (ln) {
#pragma unroll
                        for (int i = 0; i < BN; ++i) v[i] = mr.y * fmaf(-mr.x, lnc[i], v[i]);
                    }
#pragma unroll
                    for (int i = 0; i < BN; ++i) v[i] += add[i];
                    if (mode == EPI_RESADD) {
                        float* o = reinterpret_cast<float*>(ep.out) + row;
#pragma unroll
                        for (int i = 0; i < BN; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (ep.x16_out != nullptr) {
                            // folded LayerNorm, producer side: the new residual rounded to h16 (the next GEMM's A operand) and
                            // the (mean, M2) of every 32-column piece of this row segment
                            h16* o16 = reinterpret_cast<h16*>(ep.x16_out) + row;
#pragma unroll
                            for (int i = 0; i < BN; i += 8) {
                                uint4 u;
                                u.x = pack_h16x2(v[i], v[i + 1]); u.y = pack_h16x2(v[i + 2], v[i + 3]);
                                u.z = pack_h16x2(v[i + 4], v[i + 5]); u.w = pack_h16x2(v[i + 6], v[i + 7]);
                                *reinterpret_cast<uint4*>(o16 + i) = u;
                            }
#pragma unroll
                            for (int pc = 0; pc < BN / WIPA_LN_PIECE; ++pc) {
                                const float2 st2 = ln_piece_stats(v + pc * WIPA_LN_PIECE);
                                *reinterpret_cast<float2*>(ep.ln_stats_out +
                                                           ((long long)m * (ep.N / WIPA_LN_PIECE) + blockIdx.x * (BN / WIPA_LN_PIECE) + pc) * 2) = st2;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < BN; i += 8) epi_group<8, true, MODE>(ep, m, n0 + i, v + i);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < BN; i += 8) epi_group<8, false, MODE>(ep, m, n0 + i, v + i);
                }
            }
        } else {
        // the tile's bias goes to shared memory while the main loop runs (weights are static: no dependency), so the
        // column chunks below never wait on a global load
        __shared__ float s_bias[BN];
        __shared__ float s_lnc[BN];
        const bool ln = ep.ln_stats != nullptr;
        const bool have_bias = ep.bias != nullptr && (mode != EPI_ARGMAX || ln);
        if (have_bias || ln) {
            for (int i = threadIdx.x - 64; i < BN; i += 128) {
                s_bias[i] = (have_bias && n0 + i < ep.N) ? ep.bias[wn0 + i] : 0.f;                   // indexed like the rows of W
                s_lnc[i] = (ln && n0 + i < ep.N) ? ep.ln_c[n0 + i] : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const float* bias_tile = (have_bias && mode != EPI_ARGMAX) ? s_bias : nullptr;
        // folded LayerNorm, consumer side (see the 32-column path): y = rstd * (acc - mean * c[n]) (+ bias' in the epilogue)
        float2 mr = make_float2(0.f, 1.f);
        if (ln && row_ok && !(BOXM == 64 && quarter >= 2)) mr = ln_row_stats(ep.ln_stats + (long long)m * ep.ln_nt * 2, ep.ln_nt);
        ptx::mbar_wait(tmem_full, 0);
        if (threadIdx.x == 64) DBG_STAMP(5);
        ptx::tc_fence_after();
        if (BOXM == 64 && quarter >= 2) {
            // lanes 64..127 hold garbage (see TcCfg); nothing to do
        } else if (mode == EPI_ARGMAX) {
            const bool begin = (ep.step_ptr != nullptr) && (*ep.step_ptr == 0);
            float best = -INFINITY;
            int best_n = 0x7fffffff;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                ptx::tmem_ld16(taddr + c0, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = n0 + c0 + i;
                    const float x = ln ? mr.y * fmaf(-mr.x, s_lnc[c0 + i], v[i]) + s_bias[c0 + i] : v[i];
                    if (n < ep.N && !vocab_masked(ep, n, begin) && x > best) { best = x; best_n = n; }
                }
            }
            if (row_ok) {
                ep.pmax[(long long)m * ep.n_tiles + blockIdx.x] = best;
                ep.pidx[(long long)m * ep.n_tiles + blockIdx.x] = best_n;
            }
        } else {
#pragma unroll 2
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                ptx::tmem_ld16(taddr + c0, v);
                ptx::tmem_ld_wait();
                if (ln) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = mr.y * fmaf(-mr.x, s_lnc[c0 + i], v[i]);
                }
                if (row_ok) {
                    epi_group<8, false, MODE>(ep, m, n0 + c0, v, bias_tile, n0);
                    epi_group<8, false, MODE>(ep, m, n0 + c0 + 8, v + 8, bias_tile, n0);
                }
            }
        }
        }
    }
    if (threadIdx.x == 64) DBG_STAMP(6);
    ptx::tc_fence_before();
    __syncthreads();
    if (CL > 1 && (num_kb + Cfg::KPB - 1) / Cfg::KPB > Cfg::STAGES) ptx::cluster_sync_all();   // no CTA leaves while peers may still signal its barriers
    if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    if (threadIdx.x == 0) DBG_STAMP(7);
}

int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(map, WIPA_H16_TMA_TYPE, (cuuint32_t)rank, const_cast<void*>(base), dims,
                          strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wipa_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu strides %llu,%llu box %u,%u", (int)r,
                       rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                       (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0],
                       (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], box[1]);
        return WIPA_ECUDA;
    }
    return WIPA_OK;
}

template <int BN, int BOXM, int CL, int MODE>
int launch_bn_mode(const CUtensorMap& tmA, const CUtensorMap& tmW, int num_kb, int tiles_per_batch, int n_batch, int a_rpb,
                   int N, EpiParams ep, cudaStream_t st, int splits) {
    using Cfg = TcCfg<BN, BOXM>;
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(gemm_h16_tc_kernel<BN, BOXM, CL, MODE>, (size_t)Cfg::SMEM, attr));
    dim3 grid(cdiv(N, BN), tiles_per_batch * n_batch, splits);
    if (ep.mode == EPI_ARGMAX) ep.n_tiles = grid.x;
    if (CL == 1) {
        WIPA_CUDA_CHECK(wipa_launch(gemm_h16_tc_kernel<BN, BOXM, CL, MODE>, grid, dim3(192), (size_t)Cfg::SMEM, st, tmA, tmW, num_kb,
                                    tiles_per_batch, a_rpb, ep));
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (g_wipa_pdl & WIPA_PDL_CLASS) ? 2 : 1;
        WIPA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_h16_tc_kernel<BN, BOXM, CL, MODE>, tmA, tmW, num_kb, tiles_per_batch, a_rpb, ep));
    }
    WIPA_LAUNCHED();
    return WIPA_OK;
}

// The decode step's (tile width, epilogue) pairs get specialised instantiations; everything else (encoder shapes below the
// persistent kernel's threshold, tests, the multicast experiment) takes the run-time switch.
template <int BN, int BOXM, int CL>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmW, int num_kb, int tiles_per_batch, int n_batch, int a_rpb,
              int N, const EpiParams& ep, cudaStream_t st, int splits = 1) {
#define WIPA_TC_CASE(M) case M: return launch_bn_mode<BN, BOXM, CL, M>(tmA, tmW, num_kb, tiles_per_batch, n_batch, a_rpb, N, ep, st, splits)
    if constexpr (CL == 1 && BN == 32) {
        switch (ep.mode) { WIPA_TC_CASE(EPI_QKV_DEC); WIPA_TC_CASE(EPI_RESADD); WIPA_TC_CASE(EPI_STORE); default: break; }
    } else if constexpr (CL == 1 && BN == 64) {
        switch (ep.mode) { WIPA_TC_CASE(EPI_GELU); WIPA_TC_CASE(EPI_STORE); WIPA_TC_CASE(EPI_QKV_DEC); WIPA_TC_CASE(EPI_RESADD); default: break; }
    } else if constexpr (CL == 1 && BN == 128) {
        switch (ep.mode) { WIPA_TC_CASE(EPI_ARGMAX); WIPA_TC_CASE(EPI_STORE); WIPA_TC_CASE(EPI_GELU); default: break; }
    } else if constexpr (CL == 1 && BN == 256) {
        switch (ep.mode) { WIPA_TC_CASE(EPI_STORE); default: break; }
    }
#undef WIPA_TC_CASE
    return launch_bn_mode<BN, BOXM, CL, -1>(tmA, tmW, num_kb, tiles_per_batch, n_batch, a_rpb, N, ep, st, splits);
}

}  // namespace

int wipa_init_tma() {
    if (g_encode != nullptr) return WIPA_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    WIPA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    WIPA_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, WIPA_ECUDA,
               "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return WIPA_OK;
}

int gemm_h16_num_tiles(int N, int block_n) { return cdiv(N, block_n); }

extern "C" int wipa_debug_gemm_stamps(unsigned long long* host_out, int n) {
#ifdef WIPA_GEMM_DBG
    return cudaMemcpyFromSymbol(host_out, g_gemm_dbg, sizeof(unsigned long long) * (size_t)n) == cudaSuccess ? 0 : -2;
#else
    (void)host_out; (void)n;
    return WIPA_EUNSUPPORTED;
#endif
}

int launch_gemm_h16(const AOperand& a, const h16* W, int M, int N, int K, const EpiParams& ep_in, int block_n,
                     cudaStream_t st) {
    if (block_n == 0) return launch_gemm_h16_persistent(a, W, M, N, K, ep_in, st);     // 128 x 256 persistent kernel
    WIPA_TRY(wipa_init_tma());
    WIPA_CHECK(K % 8 == 0 && a.lda % 8 == 0 && a.a_bstride % 8 == 0, WIPA_EINVAL,
               "gemm_h16: K / lda / batch stride must be multiples of 8 elements (16 bytes)");
    WIPA_CHECK(M == a.a_rpb * a.n_batch, WIPA_EINVAL, "gemm_h16: M != rows_per_batch * batches");
    WIPA_CHECK((reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, WIPA_EINVAL,
               "gemm_h16: operands must be 16-byte aligned");
    WIPA_CHECK(ep_in.ln_stats == nullptr || block_n != 32 || (N % 32 == 0 && ep_in.vec_ok), WIPA_EINVAL,
               "gemm_h16: the folded-LayerNorm epilogue of the 32-column tiles needs N %% 32 == 0");
    WIPA_CHECK(ep_in.x16_out == nullptr || ((block_n == 32 || block_n == 64) && N % block_n == 0 && ep_in.vec_ok && ep_in.mode == EPI_RESADD), WIPA_EINVAL,
               "gemm_h16: residual statistics are produced by 32- / 64-column EPI_RESADD tiles only");
    const int box_m = (a.a_rpb <= 64 && block_n == 32) ? 64 : 128;
    // WIPA_GEMM_MULTICAST=1: clusters of 4 N tiles with A multicast for the narrow (decode) tiles whenever the N tiles
    // divide evenly.  Parity-tested, but OFF by default: measured on B200 at B=256 the two cluster barriers and the
    // co-scheduling constraint cost more (3345 us / decode step) than the saved L2->SM traffic (3204 us without).
    static const int mc_env = [] { const char* e = getenv("WIPA_GEMM_MULTICAST"); return (e && *e) ? atoi(e) : 0; }();
    const bool mc = mc_env != 0 && block_n <= 64 && (cdiv(N, block_n) % 4 == 0);
    CUtensorMap tmA, tmW;
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)a.a_rpb, (cuuint64_t)a.n_batch};
        cuuint64_t bstride = a.n_batch > 1 ? (cuuint64_t)a.a_bstride : (cuuint64_t)a.a_rpb * (cuuint64_t)a.lda;
        cuuint64_t strides[2] = {(cuuint64_t)a.lda * 2, bstride * 2};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)(mc ? box_m / 4 : box_m), 1};
        WIPA_TRY(make_map(&tmA, a.ptr, 3, dims, strides, box));
    }
    {
        // batched W: n_batch blocks of w_brows rows; the box of the last n-tile of a block may run into the next block (or past the
        // end, zero-filled): those columns are >= N and never stored
        WIPA_CHECK(ep_in.w_brows == 0 || ep_in.w_brows >= N, WIPA_EINVAL, "gemm_h16: w_brows %d < N %d", ep_in.w_brows, N);
        cuuint64_t dims[2] = {(cuuint64_t)K, ep_in.w_brows > 0 ? (cuuint64_t)ep_in.w_brows * (cuuint64_t)a.n_batch : (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {TC_BK, (cuuint32_t)block_n};
        WIPA_TRY(make_map(&tmW, W, 2, dims, strides, box));
    }
    EpiParams ep = ep_in;
    ep.M_rows = a.a_rpb;
    const int num_kb = cdiv(K, TC_BK);
    const int tpb = cdiv(a.a_rpb, TC_BM);
    // split-K (3 ways) when the caller provides scratch and K is long: the narrow tiles are bound by how fast ONE SM can
    // stream its 128 x K activations + BN x K weights, so a long K wants more CTAs, not wider ones
    int splits = 1;
    if (ep.sk_part != nullptr && num_kb >= 24 && (block_n == 32 || (block_n == 64 && ep.mode == EPI_RESADD && N % 64 == 0 && ep.vec_ok)))
        splits = ep.sk_splits > 0 ? ep.sk_splits : 3;
    if (splits > num_kb) splits = num_kb;
    if (splits == 1) { ep.sk_part = nullptr; ep.sk_count = nullptr; }
    switch (block_n) {
        case 32:
            if (box_m == 64) return mc ? launch_bn<32, 64, 4>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits)
                                       : launch_bn<32, 64, 1>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits);
            return mc ? launch_bn<32, 128, 4>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits)
                      : launch_bn<32, 128, 1>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits);
        case 64: return mc ? launch_bn<64, 128, 4>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits)
                           : launch_bn<64, 128, 1>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st, splits);
        case 128: return launch_bn<128, 128, 1>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st);
        case 256: return launch_bn<256, 128, 1>(tmA, tmW, num_kb, tpb, a.n_batch, a.a_rpb, N, ep, st);
        default: break;
    }
    wipa_set_error("gemm_h16: unsupported block_n %d", block_n);
    return WIPA_EINVAL;
}
