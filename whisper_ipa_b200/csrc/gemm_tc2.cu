// Encoder-side h16 GEMM on tcgen05: persistent, warp-specialised, double-buffered accumulators.
//
//   C[M,N] = A[M,K] * W[N,K]^T with M in the tens of thousands (32 clips x 1500 frames per encoder micro-batch).
//
//   grid        one CTA per SM; CTA i walks tiles i, i + grid, ... in n-fastest order (CTAs that run together share
//               the A tile and all of W in L2)
//   warp 0      TMA producer: A (128 x 64) and W (256 x 64) k-blocks, 128B-swizzled, 3-deep ring across tile boundaries
//   warp 1      tcgen05 issuer: M128 x N256 x K16 MMAs (128 cycles each, 96 B/cycle of operand reads - under the 128
//               B/cycle shared-memory limit that caps N = 128 tiles) into one of TWO 256-column TMEM accumulators
//   warps 2-9   epilogue, overlapped with the next tile's main loop: tcgen05.ld (lane = row), then per instantiation
//               (G2_EP_*): h16 results -> bias + packed-f32x2 GELU by the row owner -> h16 staging tile -> 16-byte
//               pieces, 8 lanes per 128-byte output line; fp32 results -> fp32 staging tile -> (row, 8-column) items
//               with coalesced residual loads; anything else -> the generic run-time epilogue (common.cuh epi_group)
//
// The A operand is the same 3-D tensor map as in gemm_tc.cu (rows may overlap: conv1d(k=3) as a GEMM).
#define WIPA_PDL_CLASS 8
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int G2_BM = 128;
constexpr int G2_BN = 256;
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 3;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;                 // 16 KB
constexpr int G2_W_BYTES = G2_BN * G2_BK * 2;                 // 32 KB
constexpr int G2_LDS = 68;                                    // staging row pitch in floats (64 + 4: conflict-free float4 rows)
constexpr int G2_STAGE_F = G2_BM * G2_LDS;                    // floats per staging tile (128 rows x 64 columns)
constexpr int G2_SMEM = G2_STAGES * (G2_A_BYTES + G2_W_BYTES) + 2 * G2_STAGE_F * 4 + 1024 /*align*/ + 256 /*barriers*/ +
                        2048 /*bias tile + folded-LayerNorm column sums*/;
constexpr int G2_THREADS = 64 + 256;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g2_encode = nullptr;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Epilogue variants.  The encoder's four shapes get their own instantiation so that each carries only the code it runs
// (the generic kernel's SASS is ~64 KB and the eight epilogue warps sit at different places in it: 6 % of its samples
// were instruction-fetch stalls); anything else takes G2_EP_GENERIC, which decides per chunk at run time.
enum { G2_EP_GENERIC = 0, G2_EP_HEADS_H16 = 1, G2_EP_RESADD_F32 = 2, G2_EP_GELU_H16 = 3, G2_EP_STORE_H16 = 4, G2_EP_ARGMAX = 5 };

// Two GELUs at a time on the packed fp32 pipe (fma/mul.f32x2): 9 issue slots per value instead of 15.  Same A&S 7.1.26
// erf as gelu_erf_fast; the sign select is folded into gelu(x) = max(x, 0) - |x| * h(|x|).
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_dup(float a) { return f2_pack(a, a); }
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
    float t0, t1, e0, e1, a0, a1, h0, h1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.23164189f, fabsf(x0), 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.23164189f, fabsf(x1), 1.0f)));
    const uint64_t t = f2_pack(t0, t1), x = f2_pack(x0, x1);
    uint64_t p = f2_fma(f2_dup(0.5307027145f), t, f2_dup(-0.7265760135f));
    p = f2_fma(p, t, f2_dup(0.7107068705f));
    p = f2_fma(p, t, f2_dup(-0.142248368f));
    p = f2_fma(p, t, f2_dup(0.127414796f));
    f2_unpack(f2_mul(f2_mul(x, f2_dup(-0.72134752f)), x), a0, a1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    f2_unpack(f2_mul(p, f2_mul(t, f2_pack(e0, e1))), h0, h1);
    x0 = fmaf(-fabsf(x0), h0, fmaxf(x0, 0.f));
    x1 = fmaf(-fabsf(x1), h1, fmaxf(x1, 0.f));
}

template <int EP>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_h16_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int num_kb,
                            int tiles_per_batch, int n_mtiles, int n_ntiles, int a_rpb, EpiParams ep) {
    extern __shared__ uint8_t g2_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g2_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = sA + G2_STAGES * G2_A_BYTES;
    float* stage_f = reinterpret_cast<float*>(sW + G2_STAGES * G2_W_BYTES);       // [2 halves][128][G2_LDS]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_f + 2 * G2_STAGE_F);
    uint64_t* empty = full + G2_STAGES;
    uint64_t* tmem_full = empty + G2_STAGES;                  // [2]
    uint64_t* tmem_empty = tmem_full + 2;                     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(full) + 64;          // [2 halves][128], after the 256 bytes of barriers
    float* s_lnc = s_bias + 256;                                  // [2 halves][128] (folded LayerNorm: column sums of W')

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = n_mtiles * n_ntiles;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < G2_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
            for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], 8); }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch (common.cuh): barriers and TMEM are set up; everything below touches activations of
    // earlier kernels, so wait for them here, and only then let the next kernel begin its own prologue
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mt = tile / n_ntiles, nt = tile - mt * n_ntiles;
            const int batch = mt / tiles_per_batch;
            const int t0 = (mt - batch * tiles_per_batch) * G2_BM;
            const int n0 = nt * G2_BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&empty[s], ph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&full[s], G2_A_BYTES + G2_W_BYTES);
                    ptx::tma_load_3d(sA + s * G2_A_BYTES, &tmA, &full[s], kb * G2_BK, t0, batch);
                    ptx::tma_load_2d(sW + s * G2_W_BYTES, &tmW, &full[s], kb * G2_BK, n0);
                }
                __syncwarp();
                if (++s == G2_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = ptx::idesc_h16_f32(G2_BM, G2_BN);
        const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA));
        const uint32_t w_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sW));
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
            ptx::tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)buf * G2_BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)s * (G2_A_BYTES >> 4);
                    const uint32_t w_lo = w_lo0 + (uint32_t)s * (G2_W_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < G2_BK / 16; ++k)
                        ptx::umma_h16(acc, ptx::smem_desc_sw128(a_lo + 2 * k), ptx::smem_desc_sw128(w_lo + 2 * k), idesc,
                                       (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == G2_STAGES) { s = 0; ph ^= 1; }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tmem_full[buf]);
            __syncwarp();
        }
    } else {
        const int ew = warp - 2;                               // 0..7
        const int half = ew >> 2;                              // which 128 columns of the accumulator
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;                     // row of the tile held by this thread after tcgen05.ld
        const int tih = (ew & 3) * 32 + lane;                  // thread index within the half-group (0..127)
        float* st = stage_f + half * G2_STAGE_F;
        const uint32_t st_s = ptx::smem_u32(st);                // shared-space address of this half-group's staging tile
        const int bar_id = 1 + half;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / n_ntiles, nt = tile - mt * n_ntiles;
            const int batch = mt / tiles_per_batch;
            const int t0 = (mt - batch * tiles_per_batch) * G2_BM;
            const int n0 = nt * G2_BN + half * 128;
            const int buf = it & 1;
            constexpr bool kBf16Staged = EP == G2_EP_HEADS_H16 || EP == G2_EP_GELU_H16 || EP == G2_EP_STORE_H16;
            float bias_mine = 0.f;                              // column n0 + tih of the bias, fetched ahead of the accumulator
            if (kBf16Staged && ep.bias != nullptr && n0 + tih < ep.N) bias_mine = __ldg(ep.bias + n0 + tih);
            ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            if (kBf16Staged) s_bias[half * 128 + tih] = bias_mine;   // read after the next named barrier; the previous tile's
                                                                   // readers are behind its last one
            const uint32_t taddr = tmem_base + (uint32_t)buf * G2_BN + (uint32_t)half * 128u + ((uint32_t)(quarter * 32) << 16);
            if constexpr (EP == G2_EP_ARGMAX) {
                // fused vocabulary argmax (decode step): this thread's row, the 128 columns of its half; nothing is stored but
                // one (max, argmax) pair per (row, 128-column piece).  Suppress masks arrive as two 32-bit words per 64 columns.
                const bool begin = (ep.step_ptr != nullptr) && (*ep.step_ptr == 0);
                float best = -INFINITY;
                int best_n = 0x7fffffff;
                // folded final LayerNorm (common.cuh): logits = rstd * (acc - mean * c[n]) + b'[n]; this half's 128 values of
                // c and b' go through shared memory, the row's (mean, rstd) comes from the producer's per-piece statistics
                const bool ln = ep.ln_stats != nullptr;
                float2 mr = make_float2(0.f, 1.f);
                if (ln) {
                    named_bar_sync(bar_id, 128);                   // the previous tile's readers are done with s_bias / s_lnc
                    s_bias[half * 128 + tih] = (n0 + tih < ep.N) ? __ldg(ep.bias + n0 + tih) : 0.f;
                    s_lnc[half * 128 + tih] = (n0 + tih < ep.N) ? __ldg(ep.ln_c + n0 + tih) : 0.f;
                    if (t0 + r < ep.M_rows) mr = ln_row_stats(ep.ln_stats + ((long long)batch * a_rpb + t0 + r) * ep.ln_nt * 2, ep.ln_nt);
                    named_bar_sync(bar_id, 128);
                }
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[64];
                    ptx::tmem_ld32(taddr + c * 64, v);
                    ptx::tmem_ld32(taddr + c * 64 + 32, v + 32);
                    ptx::tmem_ld_wait();
                    const int nc0 = n0 + c * 64;
                    if (ln) {
                        const uint32_t sb = ptx::smem_u32(s_bias + half * 128 + c * 64), sc = ptx::smem_u32(s_lnc + half * 128 + c * 64);
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            const float4 b4 = ptx::lds128(sb + i * 4), c4 = ptx::lds128(sc + i * 4);
                            v[i] = fmaf(mr.y, fmaf(-mr.x, c4.x, v[i]), b4.x);
                            v[i + 1] = fmaf(mr.y, fmaf(-mr.x, c4.y, v[i + 1]), b4.y);
                            v[i + 2] = fmaf(mr.y, fmaf(-mr.x, c4.z, v[i + 2]), b4.z);
                            v[i + 3] = fmaf(mr.y, fmaf(-mr.x, c4.w, v[i + 3]), b4.w);
                        }
                    }
#pragma unroll
                    for (int wd = 0; wd < 2; ++wd) {
                        const int nw = nc0 + wd * 32;              // a multiple of 32: one mask word
                        uint32_t mk = 0u;
                        if (nw < ep.N) {
                            if (ep.mask_always != nullptr) mk = __ldg(ep.mask_always + (nw >> 5));
                            if (begin && ep.mask_begin != nullptr) mk |= __ldg(ep.mask_begin + (nw >> 5));
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int n = nw + i;
                            const float x = v[wd * 32 + i];
                            if (n < ep.N && !((mk >> i) & 1u) && x > best) { best = x; best_n = n; }
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
                if (t0 + r < ep.M_rows) {
                    const long long m = (long long)batch * a_rpb + t0 + r;
                    ep.pmax[m * ep.n_tiles + nt * 2 + half] = best;
                    ep.pidx[m * ep.n_tiles + nt * 2 + half] = best_n;
                }
                continue;
            } else {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {                      // 64 columns at a time
                float v[64];
                ptx::tmem_ld32(taddr + c * 64, v);
                ptx::tmem_ld32(taddr + c * 64 + 32, v + 32);
                ptx::tmem_ld_wait();
                if (c == 1) {                                  // accumulator fully read by this warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
                }
                // ---- h16 results: bias / GELU by the row-owning thread, then a h16 staging tile ------------------------------
                // (The epilogue is issue-bound, not bandwidth-bound: ~24 % of each epilogue warp's samples are `selected`.
                // Staging the finished h16 values halves the shared-memory instructions; unstaged row-per-thread stores
                // were tried and are slower: 32 partial lines per store instruction.)
                const int nc0 = n0 + c * 64;
                if (EP != G2_EP_GENERIC && nc0 >= ep.N) continue;   // N % 64 == 0: a chunk of the last n-tile is all in or all out
                                                                    // (uniform over the half-group, so its barriers stay matched)
                if constexpr (kBf16Staged) {
                    constexpr uint32_t PITCH = 144;                // bytes per staged row: 64 h16 + 16 (conflict-free 16-byte rows)
                    named_bar_sync(bar_id, 128);               // previous chunk's readers are done with the staging tile; bias tile visible
                    if (ep.bias != nullptr) {
                        const uint32_t sb = ptx::smem_u32(s_bias + half * 128 + c * 64);
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            const float4 b4 = ptx::lds128(sb + i * 4);
                            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                        }
                    }
                    if constexpr (EP == G2_EP_GELU_H16) {
#pragma unroll
                        for (int i = 0; i < 64; i += 2) gelu_erf_fast2(v[i], v[i + 1]);
                    }
#pragma unroll
                    for (int i = 0; i < 64; i += 8) {
                        uint4 u;
                        u.x = pack_h16x2(v[i], v[i + 1]); u.y = pack_h16x2(v[i + 2], v[i + 3]);
                        u.z = pack_h16x2(v[i + 4], v[i + 5]); u.w = pack_h16x2(v[i + 6], v[i + 7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_s + (uint32_t)r * PITCH + (uint32_t)i * 2u),
                                     "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
                    }
                    named_bar_sync(bar_id, 128);
                    // readers: thread -> 16-byte piece p8 of rows rr0, rr0 + 16, ...: 8 lanes cover one 128-byte output line
                    const int p8 = tih & 7, rr0 = tih >> 3;
                    const int m_first = batch * a_rpb + t0;
                    long long base_off, wrap_off, row_pitch;
                    int row0, period;                              // row of tile row 0 within its clip / output batch, and the period
                    if constexpr (EP == G2_EP_HEADS_H16) {
                        const int which = nc0 / ep.d;
                        const int h = (nc0 - which * ep.d) >> 6;
                        const int b0 = m_first / ep.T;
                        row0 = m_first - b0 * ep.T; period = ep.T;
                        base_off = (long long)which * ep.which_stride + ((long long)(b0 * ep.H + h) * ep.T) * WIPA_HEAD_DIM + p8 * 8;
                        wrap_off = (long long)(ep.H - 1) * ep.T * WIPA_HEAD_DIM;      // next clip: + H*T rows, - T rows
                        row_pitch = WIPA_HEAD_DIM;
                    } else {
                        const int ob0 = m_first / ep.o_rpb;
                        row0 = m_first - ob0 * ep.o_rpb; period = ep.o_rpb;
                        base_off = (long long)ob0 * ep.o_bstride + nc0 + p8 * 8;
                        wrap_off = ep.o_bstride - (long long)ep.o_rpb * ep.ldo;
                        row_pitch = ep.ldo;
                    }
                    h16* const out_b = reinterpret_cast<h16*>(ep.out) + base_off;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rr = rr0 + j * 16;
                        if (t0 + rr < ep.M_rows) {
                            uint4 u;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                         : "r"(st_s + (uint32_t)rr * PITCH + (uint32_t)p8 * 16u) : "memory");
                            const int tt = row0 + rr;
                            const long long off = (long long)tt * row_pitch + (tt >= period ? wrap_off : 0ll);
                            *reinterpret_cast<uint4*>(out_b + off) = u;
                        }
                    }
                    continue;
                }
                named_bar_sync(bar_id, 128);                   // previous chunk's readers are done with the staging tile
                const uint32_t row_s = st_s + (uint32_t)(r * G2_LDS) * 4u;
#pragma unroll
                for (int i = 0; i < 64; i += 4) ptx::sts128(row_s + i * 4, v[i], v[i + 1], v[i + 2], v[i + 3]);
                named_bar_sync(bar_id, 128);
                // each thread owns the 8-column group c8 of rows rr0, rr0 + 16, ..., rr0 + 112 of this 64-column chunk
                const int c8 = tih & 7, rr0 = tih >> 3;
                const int nc = n0 + c * 64;                        // first column of the chunk
                const int ncol = nc + c8 * 8;
                const bool fast = EP == G2_EP_RESADD_F32 ||
                                  (ep.vec_ok && (nc + 64 <= ep.N) &&
                                   ((ep.mode == EPI_HEADS && ep.T >= G2_BM && ep.vt_which < 0) ||
                                    ((ep.mode == EPI_RESADD || ep.mode == EPI_GELU || ep.mode == EPI_STORE) && ep.o_rpb >= G2_BM)));
                if (fast) {
                    float bias8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) bias8[i] = 0.f;
                    if (ep.bias != nullptr) {
                        const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + ncol);
                        const float4 b1 = *reinterpret_cast<const float4*>(ep.bias + ncol + 4);
                        bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
                        bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
                    }
                    const int m_first = batch * a_rpb + t0;        // logical row of the tile's first row
                    if (EP != G2_EP_RESADD_F32 && ep.mode == EPI_HEADS) {
                        // one 64-column chunk = one head of one of q|k|v: a single division pair per chunk, and the clip
                        // index advances at most once inside a 128-row tile (T >= 128)
                        const int which = nc / ep.d;
                        const int h = (nc - which * ep.d) >> 6;
                        const int b0 = m_first / ep.T, tt0 = m_first - b0 * ep.T;
                        const long long wbase = (long long)which * ep.which_stride + (long long)h * ep.T * WIPA_HEAD_DIM + c8 * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int rr = rr0 + j * 16;
                            if (t0 + rr < ep.M_rows) {
                                int tt = tt0 + rr, bb = b0;
                                if (tt >= ep.T) { tt -= ep.T; ++bb; }
                                const float4 x0 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8) * 4u);
                                const float4 x1 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8 + 4) * 4u);
                                float w[8] = {x0.x + bias8[0], x0.y + bias8[1], x0.z + bias8[2], x0.w + bias8[3],
                                              x1.x + bias8[4], x1.y + bias8[5], x1.z + bias8[6], x1.w + bias8[7]};
                                store_group<8>(ep.out, ep.out_h16, wbase + ((long long)bb * ep.H * ep.T + tt) * WIPA_HEAD_DIM, w, true);
                            }
                        }
                    } else {
                        const int ob0 = m_first / ep.o_rpb, ot0 = m_first - ob0 * ep.o_rpb;
                        long long rowoff[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            int ot = ot0 + rr0 + j * 16, ob = ob0;
                            if (ot >= ep.o_rpb) { ot -= ep.o_rpb; ++ob; }
                            rowoff[j] = (long long)ob * ep.o_bstride + (long long)ot * ep.ldo + ncol;
                        }
                        if (EP == G2_EP_RESADD_F32 || ep.mode == EPI_RESADD) {
                            // all residual loads of the chunk are issued before the first use (one HBM round trip per chunk)
                            float4 r0[8], r1[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (t0 + rr0 + j * 16 < ep.M_rows) {
                                    r0[j] = *reinterpret_cast<const float4*>(ep.resid + rowoff[j]);
                                    r1[j] = *reinterpret_cast<const float4*>(ep.resid + rowoff[j] + 4);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int rr = rr0 + j * 16;
                                if (t0 + rr < ep.M_rows) {
                                    const float4 x0 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8) * 4u);
                                    const float4 x1 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8 + 4) * 4u);
                                    float* o = reinterpret_cast<float*>(ep.out) + rowoff[j];
                                    *reinterpret_cast<float4*>(o) = make_float4(x0.x + bias8[0] + r0[j].x, x0.y + bias8[1] + r0[j].y,
                                                                                x0.z + bias8[2] + r0[j].z, x0.w + bias8[3] + r0[j].w);
                                    *reinterpret_cast<float4*>(o + 4) = make_float4(x1.x + bias8[4] + r1[j].x, x1.y + bias8[5] + r1[j].y,
                                                                                    x1.z + bias8[6] + r1[j].z, x1.w + bias8[7] + r1[j].w);
                                }
                            }
                        } else {                                    // EPI_STORE / EPI_GELU
                            const bool gelu = ep.mode == EPI_GELU;
                            const bool gelu_fast = gelu && ep.out_h16;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int rr = rr0 + j * 16;
                                if (t0 + rr < ep.M_rows) {
                                    const float4 x0 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8) * 4u);
                                    const float4 x1 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8 + 4) * 4u);
                                    float w[8] = {x0.x + bias8[0], x0.y + bias8[1], x0.z + bias8[2], x0.w + bias8[3],
                                                  x1.x + bias8[4], x1.y + bias8[5], x1.z + bias8[6], x1.w + bias8[7]};
                                    if (gelu_fast) {                 // h16 output: two MUFU + ~14 FMA-pipe instructions per value
#pragma unroll
                                        for (int i = 0; i < 8; ++i) w[i] = gelu_erf_fast(w[i]);
                                    } else if (gelu) {
#pragma unroll
                                        for (int i = 0; i < 8; ++i) w[i] = gelu_erf(w[i]);
                                    }
                                    store_group<8>(ep.out, ep.out_h16, rowoff[j], w, true);
                                }
                            }
                        }
                    }
                } else {
#pragma unroll 2
                    for (int j = 0; j < 8; ++j) {
                        const int rr = rr0 + j * 16;
                        const int t = t0 + rr;
                        if (t < ep.M_rows) {
                            float w[8];
                            const float4 x0 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8) * 4u);
                            const float4 x1 = ptx::lds128(st_s + (uint32_t)(rr * G2_LDS + c8 * 8 + 4) * 4u);
                            w[0] = x0.x; w[1] = x0.y; w[2] = x0.z; w[3] = x0.w; w[4] = x1.x; w[5] = x1.y; w[6] = x1.z; w[7] = x1.w;
                            epi_group<8>(ep, batch * a_rpb + t, ncol, w);
                        }
                    }
                }
            }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

int g2_make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g2_encode(map, WIPA_H16_TMA_TYPE, (cuuint32_t)rank, const_cast<void*>(base), dims,
                           strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wipa_set_error("gemm_h16_persistent: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return WIPA_ECUDA;
    }
    return WIPA_OK;
}

}  // namespace

int launch_gemm_h16_persistent(const AOperand& a, const h16* W, int M, int N, int K, const EpiParams& ep_in, cudaStream_t st) {
    if (g2_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WIPA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WIPA_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, WIPA_ECUDA, "cuTensorMapEncodeTiled not available");
        g2_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    WIPA_CHECK(K % 8 == 0 && a.lda % 8 == 0 && a.a_bstride % 8 == 0, WIPA_EINVAL,
               "gemm_h16: K / lda / batch stride must be multiples of 8 elements (16 bytes)");
    WIPA_CHECK(M == a.a_rpb * a.n_batch, WIPA_EINVAL, "gemm_h16: M != rows_per_batch * batches");
    WIPA_CHECK(ep_in.mode != EPI_QKV_DEC, WIPA_EINVAL, "gemm_h16_persistent: decode-only epilogue");
    WIPA_CHECK((ep_in.ln_stats == nullptr || ep_in.mode == EPI_ARGMAX) && ep_in.x16_out == nullptr, WIPA_EINVAL,
               "gemm_h16_persistent: the folded-LayerNorm epilogue exists for the vocabulary argmax only");
    CUtensorMap tmA, tmW;
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)a.a_rpb, (cuuint64_t)a.n_batch};
        cuuint64_t bstride = a.n_batch > 1 ? (cuuint64_t)a.a_bstride : (cuuint64_t)a.a_rpb * (cuuint64_t)a.lda;
        cuuint64_t strides[2] = {(cuuint64_t)a.lda * 2, bstride * 2};
        cuuint32_t box[3] = {G2_BK, G2_BM, 1};
        WIPA_TRY(g2_make_map(&tmA, a.ptr, 3, dims, strides, box));
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {G2_BK, G2_BN};
        WIPA_TRY(g2_make_map(&tmW, W, 2, dims, strides, box));
    }
    EpiParams ep = ep_in;
    ep.M_rows = a.a_rpb;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        WIPA_CUDA_CHECK(cudaGetDevice(&dev));
        WIPA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int tpb = cdiv(a.a_rpb, G2_BM);
    const int n_mtiles = tpb * a.n_batch, n_ntiles = cdiv(N, G2_BN);
    const int n_tiles = n_mtiles * n_ntiles;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    // specialised epilogues: every 64-column chunk is full and a 128-row tile crosses at most one clip / batch boundary
    int variant = G2_EP_GENERIC;
    if (ep.vec_ok && N % 64 == 0) {
        if (ep.mode == EPI_HEADS && ep.out_h16 && ep.T >= G2_BM && ep.vt_which < 0 && ep.d % 64 == 0) variant = G2_EP_HEADS_H16;
        else if (ep.mode == EPI_RESADD && !ep.out_h16 && ep.o_rpb >= G2_BM) variant = G2_EP_RESADD_F32;
        else if (ep.mode == EPI_GELU && ep.out_h16 && ep.o_rpb >= G2_BM) variant = G2_EP_GELU_H16;
        else if (ep.mode == EPI_STORE && ep.out_h16 && ep.o_rpb >= G2_BM) variant = G2_EP_STORE_H16;
    }
    static const char* force = getenv("WIPA_G2_GENERIC");      // experiment / test switch: always the run-time epilogue
    if (force != nullptr && force[0] == '1') variant = G2_EP_GENERIC;
    if (ep.mode == EPI_ARGMAX) { variant = G2_EP_ARGMAX; ep.n_tiles = 2 * n_ntiles; }     // one (max, argmax) per 128-column piece
#define G2_LAUNCH(EPV)                                                                                                        \
    {                                                                                                                         \
        static SmemAttr attr;                                                                                                 \
        WIPA_TRY(wipa_ensure_smem(gemm_h16_persistent_kernel<EPV>, (size_t)G2_SMEM, attr));                                  \
        WIPA_CUDA_CHECK(wipa_launch(gemm_h16_persistent_kernel<EPV>, dim3(grid), dim3(G2_THREADS), (size_t)G2_SMEM, st, tmA, tmW,  \
                                    cdiv(K, G2_BK), tpb, n_mtiles, n_ntiles, a.a_rpb, ep));                                   \
    }
    switch (variant) {
        case G2_EP_HEADS_H16: G2_LAUNCH(G2_EP_HEADS_H16) break;
        case G2_EP_RESADD_F32: G2_LAUNCH(G2_EP_RESADD_F32) break;
        case G2_EP_GELU_H16: G2_LAUNCH(G2_EP_GELU_H16) break;
        case G2_EP_STORE_H16: G2_LAUNCH(G2_EP_STORE_H16) break;
        case G2_EP_ARGMAX: G2_LAUNCH(G2_EP_ARGMAX) break;
        default: G2_LAUNCH(G2_EP_GENERIC) break;
    }
#undef G2_LAUNCH
    WIPA_LAUNCHED();
    return WIPA_OK;
}
