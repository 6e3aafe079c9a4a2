// Shared declarations for libwipa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/wipa.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libwipa is written for sm_100a only"
#endif

// h16: the 16-bit operand / storage type of the half-precision path.  The library is compiled once per type:
// libwipa.so with IEEE fp16 (11-bit significand: logits within 1e-3 of the fp32 oracle, the default) and libwipa_bf16.so
// with -DWIPA_H16_BF16 (bfloat16).  Both feed the same tcgen05 kind::f16 / mma.sync pipelines at the same rate; only the
// conversions, the instruction-descriptor format bits, the TMA element type and the mma.sync type suffix differ.
#ifdef WIPA_H16_BF16
typedef __nv_bfloat16 h16;
typedef __nv_bfloat162 h16x2;
#define WIPA_H16_DTYPE WIPA_DTYPE_BF16
#define WIPA_H16_NAME "bf16"
#define WIPA_H16_MMA_SUFFIX "bf16.bf16"
#define WIPA_H16_ONE_X2 0x3f803f80u                    /* two 1.0 values */
#define WIPA_H16_TMA_TYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define WIPA_H16_IDESC_FMT 1u                          /* tcgen05 kind::f16 a_format / b_format: 0 = f16, 1 = bf16 */
__host__ __device__ __forceinline__ float h16_to_f32(h16 v) { return __bfloat162float(v); }
__host__ __device__ __forceinline__ h16 f32_to_h16(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float2 h16x2_to_f2(h16x2 v) { return __bfloat1622float2(v); }
__device__ __forceinline__ h16x2 f2_to_h16x2(float a, float b) { return __floats2bfloat162_rn(a, b); }
#else
typedef __half h16;
typedef __half2 h16x2;
#define WIPA_H16_DTYPE WIPA_DTYPE_F16
#define WIPA_H16_NAME "f16"
#define WIPA_H16_MMA_SUFFIX "f16.f16"
#define WIPA_H16_ONE_X2 0x3c003c00u
#define WIPA_H16_TMA_TYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define WIPA_H16_IDESC_FMT 0u
__host__ __device__ __forceinline__ float h16_to_f32(h16 v) { return __half2float(v); }
__host__ __device__ __forceinline__ h16 f32_to_h16(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float2 h16x2_to_f2(h16x2 v) { return __half22float2(v); }
__device__ __forceinline__ h16x2 f2_to_h16x2(float a, float b) { return __floats2half2_rn(a, b); }
#endif

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void wipa_set_error(const char* fmt, ...);
extern int64_t g_wipa_launches;

#define WIPA_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            wipa_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cudaGetLastError(); /* clear the (non-sticky) error so later calls are not blamed */  \
            return WIPA_ECUDA;                                                                  \
        }                                                                                       \
    } while (0)

#define WIPA_CHECK(cond, code, ...)                   \
    do {                                              \
        if (!(cond)) {                                \
            wipa_set_error(__VA_ARGS__);              \
            return (code);                            \
        }                                             \
    } while (0)

#define WIPA_TRY(expr)              \
    do {                            \
        int _r = (expr);            \
        if (_r != WIPA_OK) return _r; \
    } while (0)

// every kernel launch of the library goes through this so bench.py can report "gpu_launches"
#define WIPA_LAUNCHED()                                                                        \
    do {                                                                                       \
        ++g_wipa_launches;                                                                     \
        cudaError_t _e = cudaPeekAtLastError();                                                \
        if (_e != cudaSuccess) {                                                               \
            wipa_set_error("%s:%d: launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            cudaGetLastError();                                                                \
            return WIPA_ECUDA;                                                                 \
        }                                                                                      \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel into > 48 KB of dynamic shared memory.  The attribute is per device, so the cache is per device too
// (a process normally drives one GPU, but nothing here assumes it).
struct SmemAttr { size_t done[32] = {0}; };
template <typename K>
static inline int wipa_ensure_smem(K kernel, size_t bytes, SmemAttr& a) {
    int dev = 0;
    WIPA_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 32 && a.done[dev] >= bytes) return WIPA_OK;
    WIPA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (dev >= 0 && dev < 32) a.done[dev] = bytes;
    return WIPA_OK;
}

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): kernels of the decode step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, signal `launch_dependents` as soon as they start, prefetch
// their STATIC operands (weights, cached encoder K/V) and only then `griddepcontrol.wait` for the previous kernel.
// Launch latency, prologues and the first HBM loads of kernel N+1 therefore overlap the tail of kernel N.
// RULES: (1) every kernel launched through wipa_launch must execute pdl_wait() before it reads anything a previous
// kernel wrote and before it writes anything a previous kernel may still read; (2) launch_dependents is issued only
// AFTER the kernel's own pdl_wait() has returned, so at most two consecutive grids are ever in flight.  Triggering
// at kernel entry lets an unbounded chain of grids pile up behind one another; on B200 (driver 580) a chain of
// layernorm -> GEMM -> self-attention launched that way read stale q (measured: wrong logits with all three
// attributes on, exact results when any one of them was off).
// ------------------------------------------------------------------------------------------------
extern int g_wipa_pdl;       // 1 (default) or 0 (env WIPA_PDL=0): attach the PDL attribute to launches
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cls: kernel family bit (1 elementwise, 2 self-attention, 4 cross-attention, 8 tcgen05 GEMM, 16 SIMT GEMM); the
// attribute is attached when (g_wipa_pdl & cls) != 0  (WIPA_PDL = bitmask, default all families)
template <typename... KArgs, typename... Args>
static inline cudaError_t wipa_launch_c(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_wipa_pdl & cls) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifndef WIPA_PDL_CLASS
#define WIPA_PDL_CLASS 1
#endif
#define wipa_launch(...) wipa_launch_c(WIPA_PDL_CLASS, __VA_ARGS__)
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// model constants
// ------------------------------------------------------------------------------------------------
#define WIPA_N_FFT 400
#define WIPA_HOP 160
#define WIPA_N_SAMPLES 480000
#define WIPA_N_FRAMES 3000
#define WIPA_N_FREQ 201
#define WIPA_T_ENC 1500
#define WIPA_T_ENC_PAD 1504      // V^T rows padded so TMA strides are multiples of 16 bytes
#define WIPA_HEAD_DIM 64
#define WIPA_MAX_TGT 448
#define WIPA_PAGE 16             // tokens per self-KV page

// ------------------------------------------------------------------------------------------------
// epilogue description shared by the SIMT-fp32 and tcgen05-h16 GEMM kernels
// ------------------------------------------------------------------------------------------------
enum EpiMode : int {
    EPI_STORE = 0,      // out[m, n] = acc + bias
    EPI_GELU = 1,       // out = gelu(acc + bias)
    EPI_RESADD = 2,     // out_f32 = resid + acc + bias
    EPI_GELU_POS = 3,   // out_f32 = gelu(acc + bias) + pos[m % T, n]      (conv2 + positional embedding)
    EPI_HEADS = 4,      // split N into (q|k|v) x heads, write [B,H,T,64] (or V^T [B,H,64,Tpad])
    EPI_QKV_DEC = 5,    // decoder self-attn projection: q -> f32 [M,d]; k,v -> paged cache at *pos_ptr
    EPI_ARGMAX = 6      // per-(row, n-tile) running (max, argmax) with suppress masks, nothing else stored
};

struct EpiParams {
    int mode;
    int out_h16;            // element type of out/out1/out2 (0 = f32, 1 = h16)
    int vec_ok;              // 8-wide vector stores legal (alignment + N % 8 == 0)
    int M_rows;              // valid rows per batch (rows beyond are padding of the M tile)
    int N;                   // valid columns
    const float* bias;       // [N] or nullptr
    void* out;
    void* out1;
    void* out2;
    long long ldo;           // elements between consecutive output rows
    int o_rpb;               // output rows per batch  (row m -> batch m / o_rpb, t = m % o_rpb)
    long long o_bstride;     // elements between batches of the output
    const float* resid;      // EPI_RESADD (same addressing as out)
    const float* pos;        // EPI_GELU_POS: [T, N]
    int T;                   // rows per clip (EPI_HEADS / EPI_GELU_POS)
    int H;                   // heads
    int d;                   // d_model
    long long which_stride;  // EPI_HEADS: N is split in blocks of d columns; block w is written at out + w * which_stride
    int vt_which;            // EPI_HEADS: this block (-1 = none) is written transposed, [B,H,64,Tpad]
    int Tpad;
    // EPI_QKV_DEC
    const int* pos_ptr;      // device scalar: position being written
    const int* block_table;  // [M, bt_stride] page ids
    int bt_stride;
    int n_pages;             // pages per layer pool (unused by addressing, kept for asserts)
    // EPI_ARGMAX
    float* pmax;             // [M, n_tiles]
    int* pidx;               // [M, n_tiles]
    int n_tiles;
    const uint32_t* mask_always;   // vocab bitmask, bit set = suppressed (may be nullptr)
    const uint32_t* mask_begin;    // applied when *step_ptr == 0 (may be nullptr)
    const int* step_ptr;
    // deterministic split-K (decode GEMMs with long K): grid.z CTAs per tile leave raw fp32 partial tiles in `sk_part`
    // [tile][split][128][BN]; the last one to arrive (ticket in `sk_count[tile]`, self-resetting) sums them in split order
    float* sk_part;
    int* sk_count;
    int sk_splits;           // K splits when sk_part is given (0: the kernel's default of 3 for 32-column tiles)
    int gelu_fast;           // EPI_GELU with a h16 result: A&S-erf GELU (gelu_erf_fast) instead of erff
    int w_brows;             // batched W (tcgen05 decode kernel): batch b multiplies rows [b * w_brows, b * w_brows + N) of W (and takes
                             // bias[b * w_brows + n]); 0 = one W
    // ---- LayerNorm folded around the decode GEMMs (16-bit decode path; see "folded LayerNorm" below) ----
    // consumer side: the A operand is the RAW residual stream rounded to h16 and W carries the LayerNorm gain; the epilogue
    // finishes the normalisation per row:  y = rstd * (acc - mean * ln_c[n]) + bias'[n]
    const float* ln_stats;   // [M, ln_nt] (mean, M2) of each 32-column piece of the row, or nullptr
    const float* ln_c;       // [N] row sums of the gain-folded weights
    int ln_nt;               // pieces per row (d / 32)
    // producer side (EPI_RESADD of the decode step, 32-column tiles): also leave the new residual rounded to h16 and the
    // (mean, M2) of this tile's 32 columns for the next consumer
    void* x16_out;           // h16 [M, N], or nullptr
    float* ln_stats_out;     // [M, N / 32] (mean, M2)
};

// ------------------------------------------------------------------------------------------------
// folded LayerNorm: LN(x) W^T + b = rstd * (x W'^T - mean * c) + b'   with W' = W * gain (per input column),
// c[n] = sum_k W'[n, k], b'[n] = b[n] + sum_k beta[k] W[n, k].  The kernel that produces the residual stream leaves its
// row statistics as per-tile (mean, M2) pairs (Welford pieces of 32 columns, merged exactly by Chan's formula), so no
// LayerNorm kernel sits between two GEMM nodes of a decode step.
// ------------------------------------------------------------------------------------------------
#define WIPA_LN_PIECE 32
// (mean, rstd) of one row from its `nt` pieces of WIPA_LN_PIECE columns each; eps as HF (1e-5), biased variance.
// nt is even and <= WIPA_LN_MAX_NT (d = 1280).  Every piece is fetched up front as independent 16-byte loads - ONE L2 round
// trip, issued before the accumulator is awaited; a dependent load per piece cost ~9 us per GEMM node.
#define WIPA_LN_MAX_NT 40
__device__ __forceinline__ float2 ln_row_stats(const float* __restrict__ stats_row, int nt) {
    const float4* p = reinterpret_cast<const float4*>(stats_row);
    float4 q[WIPA_LN_MAX_NT / 2];
#pragma unroll
    for (int i = 0; i < WIPA_LN_MAX_NT / 2; ++i) q[i] = (2 * i < nt) ? __ldcg(p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float ms = 0.f;
#pragma unroll
    for (int i = 0; i < WIPA_LN_MAX_NT / 2; ++i) ms += q[i].x + q[i].z;        // absent pieces contribute zeros
    const float mean = ms / (float)nt;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < WIPA_LN_MAX_NT / 2; ++i) {
        if (2 * i < nt) {
            const float d0 = q[i].x - mean, d1 = q[i].z - mean;
            m2 += (q[i].y + q[i].w) + (float)WIPA_LN_PIECE * (d0 * d0 + d1 * d1);
        }
    }
    return make_float2(mean, rsqrtf(m2 / (float)(nt * WIPA_LN_PIECE) + 1e-5f));
}
// (mean, M2) of WIPA_LN_PIECE values held by one thread
__device__ __forceinline__ float2 ln_piece_stats(const float* v) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < WIPA_LN_PIECE; ++i) s += v[i];
    const float mean = s * (1.0f / WIPA_LN_PIECE);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < WIPA_LN_PIECE; ++i) { const float dlt = v[i] - mean; m2 = fmaf(dlt, dlt, m2); }
    return make_float2(mean, m2);
}

// A-operand addressing shared by both GEMM kernels: row m of the logical [M, K] matrix lives at
//   A + (m / a_rpb) * a_bstride + (m % a_rpb) * lda        (elements)
// Rows may overlap (lda < K): that is how conv1d(k=3) over a channels-last, zero-padded signal becomes a GEMM.
struct AOperand {
    const void* ptr;
    long long lda;
    int a_rpb;               // rows per batch
    long long a_bstride;
    int n_batch;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
    // exact-erf GELU (HF activation_function="gelu", HF:models/whisper/configuration_whisper.py:140)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// GELU with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the h16 rounding of the stored value):
// two MUFU ops (rcp, ex2) + ~12 FMA-pipe instructions instead of erff's ~35.  1 + erf(z) is formed without
// cancellation on the negative side.  Used only where the result is rounded to h16 (the fp32 path calls erff).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    // gelu(x) = x * Phi(x); with h = 0.5 * (1 - erf(|x| / sqrt 2)) = 0.5 * poly(t) * exp(-x^2 / 2), t = 1 / (1 + p |x| / sqrt 2):
    // Phi(x) = h for x < 0 and 1 - h otherwise.  The 0.5, the 1/sqrt 2 and log2(e) are folded into the constants.
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, ax, 1.0f)));          // p / sqrt 2 = 0.3275911 * 0.70710678
    float poly = fmaf(0.5307027145f, t, -0.7265760135f);                                      // a5 / 2, a4 / 2
    poly = fmaf(poly, t, 0.7107068705f);
    poly = fmaf(poly, t, -0.142248368f);
    poly = fmaf(poly, t, 0.127414796f);
    poly *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.72134752f) * x));               // exp(-x^2 / 2)
    const float xh = x * (poly * e);                                                          // x * h
    return x < 0.f ? xh : x - xh;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(h16 v) { return h16_to_f32(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ h16 from_f32<h16>(float v) { return f32_to_h16(v); }

__device__ __forceinline__ uint32_t pack_h16x2(float a, float b) {
    h16x2 t = f2_to_h16x2(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// store W consecutive values starting at element index `idx` of a f32 or h16 array
template <int W>
__device__ __forceinline__ void store_group(void* base, int is_h16, long long idx, const float* v, bool vec) {
    if (is_h16) {
        h16* p = reinterpret_cast<h16*>(base) + idx;
        if (vec) {
            if (W == 8) {
                uint4 u;
                u.x = pack_h16x2(v[0], v[1]); u.y = pack_h16x2(v[2], v[3]);
                u.z = pack_h16x2(v[4], v[5]); u.w = pack_h16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(p) = u;
            } else {
                uint2 u;
                u.x = pack_h16x2(v[0], v[1]); u.y = pack_h16x2(v[2], v[3]);
                *reinterpret_cast<uint2*>(p) = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i) p[i] = f32_to_h16(v[i]);
        }
    } else {
        float* p = reinterpret_cast<float*>(base) + idx;
        if (vec) {
#pragma unroll
            for (int i = 0; i < W; i += 4)
                *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i) p[i] = v[i];
        }
    }
}

// Apply the epilogue to W (4 or 8) consecutive accumulator columns n0..n0+W-1 of logical row m.
// n0 is a multiple of W.  Columns >= ep.N are dropped.
// bias_tile: optional shared-memory copy of bias[tile_n0 .. tile_n0 + BN) (zero beyond N), indexed by column - tile_n0
// MODE >= 0 fixes the epilogue at compile time (the kernel instantiation then carries only that branch); -1 reads ep.mode
template <int W, bool SKIP_BIAS = false, int MODE = -1>
__device__ __forceinline__ void epi_group(const EpiParams& ep, int m, int n0, float* v, const float* bias_tile = nullptr,
                                          int tile_n0 = 0) {
    if (n0 >= ep.N) return;
    const bool full = (n0 + W <= ep.N);
    const bool vec = ep.vec_ok && full;
    if (!SKIP_BIAS && bias_tile != nullptr) {
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] += bias_tile[n0 - tile_n0 + i];
    } else if (!SKIP_BIAS && ep.bias != nullptr) {
        if (vec) {
#pragma unroll
            for (int i = 0; i < W; i += 4) {
                float4 b = *reinterpret_cast<const float4*>(ep.bias + n0 + i);
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i) if (n0 + i < ep.N) v[i] += ep.bias[n0 + i];
        }
    }
    const int ob = m / ep.o_rpb;
    const int ot = m - ob * ep.o_rpb;
    const long long row = (long long)ob * ep.o_bstride + (long long)ot * ep.ldo;
    switch (MODE >= 0 ? MODE : ep.mode) {
        case EPI_GELU:
            if (ep.gelu_fast) {
#pragma unroll
                for (int i = 0; i < W; ++i) v[i] = gelu_erf_fast(v[i]);
            } else {
#pragma unroll
                for (int i = 0; i < W; ++i) v[i] = gelu_erf(v[i]);
            }
            // fallthrough
        case EPI_STORE: {
            if (full) store_group<W>(ep.out, ep.out_h16, row + n0, v, vec);
            else {
                for (int i = 0; i < W && n0 + i < ep.N; ++i) store_group<1>(ep.out, ep.out_h16, row + n0 + i, v + i, false);
            }
            break;
        }
        case EPI_RESADD: {
            const float* r = ep.resid + row + n0;
            float* o = reinterpret_cast<float*>(ep.out) + row + n0;
            if (vec) {
#pragma unroll
                for (int i = 0; i < W; i += 4) {
                    float4 x = *reinterpret_cast<const float4*>(r + i);
                    *reinterpret_cast<float4*>(o + i) = make_float4(x.x + v[i], x.y + v[i + 1], x.z + v[i + 2], x.w + v[i + 3]);
                }
            } else {
                for (int i = 0; i < W && n0 + i < ep.N; ++i) o[i] = r[i] + v[i];
            }
            break;
        }
        case EPI_GELU_POS: {
            const int t = m % ep.T;
            const float* pp = ep.pos + (long long)t * ep.N + n0;
            float* o = reinterpret_cast<float*>(ep.out) + row + n0;
#pragma unroll
            for (int i = 0; i < W; ++i) if (n0 + i < ep.N) o[i] = gelu_erf(v[i]) + pp[i];
            break;
        }
        case EPI_HEADS: {
            const int which = n0 / ep.d;
            const int c = n0 - which * ep.d;
            const int h = c / WIPA_HEAD_DIM;
            const int e = c - h * WIPA_HEAD_DIM;
            const int b = m / ep.T;
            const int t = m - b * ep.T;
            const long long wbase = (long long)which * ep.which_stride;
            if (which == ep.vt_which) {
                const long long base = wbase + ((long long)(b * ep.H + h) * WIPA_HEAD_DIM + e) * ep.Tpad + t;
#pragma unroll
                for (int i = 0; i < W; ++i) store_group<1>(ep.out, ep.out_h16, base + (long long)i * ep.Tpad, v + i, false);
            } else {
                const long long base = wbase + ((long long)(b * ep.H + h) * ep.T + t) * WIPA_HEAD_DIM + e;
                store_group<W>(ep.out, ep.out_h16, base, v, true);
            }
            break;
        }
        case EPI_QKV_DEC: {
            const int which = n0 / ep.d;
            const int c = n0 - which * ep.d;
            if (which == 0) {
                store_group<W>(ep.out, 0, (long long)m * ep.d + c, v, true);
            } else {
                const int h = c / WIPA_HEAD_DIM;
                const int e = c - h * WIPA_HEAD_DIM;
                const int p = *ep.pos_ptr;
                const int page = ep.block_table[(long long)m * ep.bt_stride + p / WIPA_PAGE];
                const long long base = (((long long)page * ep.H + h) * WIPA_PAGE + (p % WIPA_PAGE)) * WIPA_HEAD_DIM + e;
                store_group<W>(which == 1 ? ep.out1 : ep.out2, ep.out_h16, base, v, true);
            }
            break;
        }
        default:
            break;
    }
}

__device__ __forceinline__ bool vocab_masked(const EpiParams& ep, int n, bool begin) {
    bool m = false;
    if (ep.mask_always != nullptr) m = (ep.mask_always[n >> 5] >> (n & 31)) & 1u;
    if (begin && ep.mask_begin != nullptr) m = m || ((ep.mask_begin[n >> 5] >> (n & 31)) & 1u);
    return m;
}

// warp-level helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// kernel launch entry points implemented across the .cu files (host side)
// ------------------------------------------------------------------------------------------------
int launch_gemm_f32(const AOperand& a, const float* W, int M, int N, int K, const EpiParams& ep, cudaStream_t st);
int launch_gemm_h16(const AOperand& a, const h16* W, int M, int N, int K, const EpiParams& ep, int block_n,
                     cudaStream_t st);
int launch_gemm_h16_persistent(const AOperand& a, const h16* W, int M, int N, int K, const EpiParams& ep, cudaStream_t st);   // gemm_tc2.cu
int wipa_init_tma();   // resolves cuTensorMapEncodeTiled through the runtime; idempotent

template <typename T>
int launch_layernorm(const float* x, const float* w, const float* b, T* out, int M, int d, cudaStream_t st);
template <typename T>
int launch_enc_attention(const T* q, const T* k, const T* v, T* out, int B, int H, int Tq, cudaStream_t st);
int launch_enc_attention_tc(const h16* q, const h16* k, const h16* v, h16* out, int B, int H, int T, cudaStream_t st);   // attn_tc.cu
template <typename T>
int launch_self_attention(const float* q, const T* kpool, const T* vpool, const int* block_table, int bt_stride,
                          const int* pos_ptr, T* out, int Bs, int H, cudaStream_t st, const int* anc_base = nullptr,
                          const int* flip_ptr = nullptr, int anc_L = 0);   // anc: beam ancestry [2][Bs][anc_L] (beam search)
// split-K cross-attention: part = f32 [Bs*H, n_split, 66] scratch, counters = int [Bs*H] (zero, self-resetting)
template <typename T>
int launch_cross_attention(const float* q, const T* k, const T* v, const int* utt_of_seq, T* out, float* part,
                           int* counters, int Bs, int H, int n_split, int kv_static, cudaStream_t st);
int cross_attention_default_split(int elem_bytes, int Bs, int H);
// attn_lat.cu: cross-attention over the encoder output itself (absorbed k / v projections), h16 only
int cross_attention_latent_supported(int H);
size_t cross_attention_latent_scratch_floats(int H, int max_seqs, int n_sm);
int cross_attention_latent_keys(int H);                    // keys per chunk of the kernel instantiation for H heads
size_t cross_attention_latent_tiled_elems(int H, int T);   // elements per utterance of the chunk-tiled encoder output
int launch_cross_attention_latent(const h16* Qp, const h16* E, int tiled, int U, const int* utt_of_seq, h16* C, int S, int H, int T,
                                  float* part, size_t part_floats, int* counters, cudaStream_t st, int beams = 1);
// Chunk-tiled layout of one utterance's encoder output E [T, d = 64 H] for the latent cross-attention kernel:
// [chunk of `keys` keys][column tile h of 64][key][64 columns], the eight 16-byte pieces of every 128-byte row permuted by
// (piece ^ (key & 7)) - byte for byte what a 128B-swizzled TMA box [keys][64] leaves in shared memory, so that a whole
// chunk (keys * d * 2 bytes, contiguous) is fetched with plain bulk copies.  Element offset of (t, col):
__host__ __device__ __forceinline__ size_t lat_tile_offset(int t, int col, int H, int keys) {
    const int ch = t / keys, key = t - ch * keys, h = col >> 6, e = col & 63;
    return ((size_t)(ch * H + h) * keys + key) * 64 + (size_t)((((e >> 3) ^ (key & 7)) << 3) | (e & 7));
}
// x f32 [rows of T per utterance, d] -> LayerNorm -> chunk-tiled h16 image of utterances u0, u0 + 1, ... (elementwise.cu)
int launch_layernorm_lat(const float* x, const float* w, const float* b, h16* out_tiled, int M, int d, int T, int u0, int keys,
                         cudaStream_t st);
// row-major [U, T, d] (f32 or h16) -> chunk-tiled h16
// 16 / 20 heads (whisper-medium / large*): attn_lat_wide.cu, 128 columns per warp, 20 heads as two CTAs of 10; chunk-tiled E only
int cross_attention_latent_wide_supported(int H);
size_t cross_attention_latent_wide_scratch_floats(int H, int max_seqs, int n_sm);
int launch_cross_attention_latent_wide(const h16* Qp, const h16* E, const int* utt_of_seq, h16* C, int S, int H, int T, float* part,
                                       size_t part_floats, int* counters, cudaStream_t st, int beams = 1);
int launch_xlq_fused(const h16* A, const h16* Wq, const h16* WkT, const float* bias, const float* ln_stats, const float* ln_c,
                     int ln_nt, h16* out, int S, int H, cudaStream_t st);
int launch_lat_tile(const void* src, int src_is_h16, h16* out_tiled, int U, int T, int H, int keys, cudaStream_t st);

// elementwise.cu
template <typename T>
int launch_mel_to_rows(const float* mel, T* rows, int B, int C, cudaStream_t st);       // [B,C,3000] -> [B,3002,C] (rows 1..3000)
template <typename T>
int launch_embed(const T* tok_emb, const float* pos_emb, const int* tok, const int* pos_ptr, float* x, int Bs, int d,
                 cudaStream_t st);
int launch_embed_lnf(const h16* tok_emb, const float* pos_emb, const int* tok, const int* pos_ptr, float* x, h16* x16, float* stats,
                     int Bs, int d, cudaStream_t st);   // folded-LayerNorm decode path: + x in h16 and per-piece (mean, M2)
int launch_convert(const float* src, void* dst, long long n, float scale, int to_h16, cudaStream_t st);
int launch_conv_weight(const float* src, void* dst, int N, int C, int to_h16, cudaStream_t st);   // [N,C,3] -> [N,3*C]
int launch_row_argmax(const float* logits, int Bs, int V, const uint32_t* mask_always, const uint32_t* mask_begin,
                      const int* step_ptr, float* pmax, int* pidx, cudaStream_t st);

struct DecodeState {       // all device pointers, owned by the context
    int* pos;              // position of the token being consumed this step
    int* step;             // index of the token being sampled (pos - (n_forced - 1)); < 0 while forcing
    int* cur_tok;          // [Bs]
    int* done;             // [Bs]
    int* n_done;           // count of finished rows
    int* ticket;           // CTA arrival counter of the finalize kernel (zero between launches)
    const int* forced;     // [Bs, n_forced] teacher-forced tokens (the prompt)
    int n_forced;
    int* out_ids;          // [Bs, max_new]
    int* out_len;          // [Bs]
    int max_new;
    int eot;
};
int launch_greedy_finalize(const float* pmax, const int* pidx, int n_tiles, DecodeState ds, int Bs, cudaStream_t st);

// beam search state (beam.cu); all device pointers owned by the context.  Sequences / ancestry are ping-pong arrays
// [2][n_utts * beams][L]; *flip selects the current half.
struct BeamState {
    int* pos;                // shared with DecodeState: position consumed by the current step
    int* step;               // index of the token being sampled
    int* flip;
    int* cur_tok;            // [n_seqs] token fed to the next step
    int* run_seq;            // running sequences (prompt included)
    int* fin_seq;            // finished hypotheses, best first
    int* anc;                // anc[s][q]: slot whose self-KV pages hold position q of running beam s
    float* run_score;        // [n_seqs]
    float* run_score_next;
    float* fin_score;        // [n_seqs] length-penalised scores of the finished slots (-1e9 = empty)
    float* fin_score_next;
    int* fin_done;           // [n_seqs]
    int* fin_done_next;
    int* unsat;              // [n_utts] HF's is_early_stop_heuristic_unsatisfied
    int* n_done;             // utterances that can no longer improve
    int beams, L, prompt_len, max_length, eot;
    float length_penalty;
};
int launch_beam_row_topk(const float* logits, long long ld, int V, const uint32_t* mask_always, const uint32_t* mask_begin,
                         const int* step_ptr, const float* run_score, int keep, float* out_val, int* out_idx, int n_rows,
                         cudaStream_t st);
int launch_beam_update(const BeamState& bs, const float* cand_val, const int* cand_idx, int V, int n_utts, cudaStream_t st);
int launch_beam_init(const BeamState& bs, const int* prompt_dev, int n_utts, cudaStream_t st);
int launch_beam_finish(const BeamState& bs, int n_utts, int max_new, int* out_ids, int* out_len, cudaStream_t st);
