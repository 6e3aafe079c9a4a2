// fp32 path GEMM: C[M,N] = A[M,K] * W[N,K]^T with true fp32 FMA (no TF32, no tensor cores), so that the
// greedy token ids of the fp32 path stay bit-exact against the fp32 oracle (the top-1/top-2
// logit margin of a random-init model is ~1e-2, any reduced-precision product flips tokens).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, register-prefetched global loads.
// Shares the epilogue (bias / GELU / residual / head split / paged-KV scatter) with the tcgen05 kernel.
#define WIPA_PDL_CLASS 16
#include "common.cuh"

#define SG_BM 64
#define SG_BN 64
#define SG_BK 16
#define SG_PAD 4

__global__ void __launch_bounds__(256)
gemm_f32_kernel(AOperand a, const float* __restrict__ W, int M, int N, int K, EpiParams ep) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Ws[2][SG_BK][SG_BN + SG_PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    pdl_wait();
    pdl_launch_dependents();                     // only after our own dependency is met: at most two grids overlap

    // global->smem mapping: each thread moves one float4 of A and one of W per k-block
    const int lr = tid >> 2;            // 0..63 row within tile
    const int lk = (tid & 3) * 4;       // 0,4,8,12
    const int am = m0 + lr;
    const float* a_row = nullptr;
    if (am < M) {
        const int b = am / a.a_rpb;
        const int t = am - b * a.a_rpb;
        a_row = reinterpret_cast<const float*>(a.ptr) + (long long)b * a.a_bstride + (long long)t * a.lda;
    }
    const int wn = n0 + lr;
    const float* w_row = wn < N ? W + (long long)wn * K : nullptr;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = K / SG_BK;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rw = ra;
    if (a_row) ra = *reinterpret_cast<const float4*>(a_row + lk);
    if (w_row) rw = *reinterpret_cast<const float4*>(w_row + lk);
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        As[buf][lk + 0][lr] = ra.x; As[buf][lk + 1][lr] = ra.y; As[buf][lk + 2][lr] = ra.z; As[buf][lk + 3][lr] = ra.w;
        Ws[buf][lk + 0][lr] = rw.x; Ws[buf][lk + 1][lr] = rw.y; Ws[buf][lk + 2][lr] = rw.z; Ws[buf][lk + 3][lr] = rw.w;
        __syncthreads();
        if (kb + 1 < nk) {
            const int ko = (kb + 1) * SG_BK + lk;
            ra = a_row ? *reinterpret_cast<const float4*>(a_row + ko) : make_float4(0.f, 0.f, 0.f, 0.f);
            rw = w_row ? *reinterpret_cast<const float4*>(w_row + ko) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
        // the next iteration writes the other buffer; one barrier per k-block is enough with 2 buffers
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m < M) epi_group<4>(ep, m, n0 + tx * 4, acc[i]);
    }
}

int launch_gemm_f32(const AOperand& a, const float* W, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
    WIPA_CHECK(K % SG_BK == 0, WIPA_EINVAL, "gemm_f32: K=%d not a multiple of %d", K, SG_BK);
    WIPA_CHECK(a.lda % 4 == 0 && a.a_bstride % 4 == 0, WIPA_EINVAL, "gemm_f32: A rows must be 16-byte aligned");
    WIPA_CHECK(ep.mode != EPI_ARGMAX, WIPA_EINVAL, "gemm_f32: fused argmax is a tcgen05-path epilogue");
    dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM));
    WIPA_CUDA_CHECK(wipa_launch(gemm_f32_kernel, grid, dim3(256), (size_t)0, st, a, W, M, N, K, ep));
    WIPA_LAUNCHED();
    return WIPA_OK;
}
