// Microbenchmark (development aid): MUFU.EX2 issue rate per SM sub-partition on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
// One CTA per SM, W warps per CTA (W / 4 per scheduler); every thread runs N independent ex2 chains of 8 registers.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ unsigned pack_h2(float a, float b) {
    unsigned r;
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// 0: ex2 only, 1: ffma + ex2 + fadd per element, 2: + one f16x2 pack per two elements, 3: + one 16-byte shared store per eight
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    __shared__ uint4 sbuf[1024];
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = -1.0f - 0.001f * (threadIdx.x + i);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2(v[i]) - 1.5f;          // keeps the argument in range; the FADD is dependent but cheap
            else { p[i] = ex2(fmaf(v[i], 1.4426950408889634f, -0.25f)); acc += p[i]; v[i] = v[i] * 0.999f - 0.001f; }
        }
        if (MODE >= 2) {
            uint4 u;
            u.x = pack_h2(p[0], p[1]); u.y = pack_h2(p[2], p[3]); u.z = pack_h2(p[4], p[5]); u.w = pack_h2(p[6], p[7]);
            if (MODE == 3) sbuf[threadIdx.x] = u;
            else acc += __uint_as_float((u.x ^ u.y ^ u.z ^ u.w) & 0x3f800000u);
        }
    }
    const long long t1 = clock64();
    float s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (MODE == 3) s += __uint_as_float(sbuf[(threadIdx.x + 1) & 1023].x & 0x3f800000u);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
                else if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
                else if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
                else k<3><<<148, warps * 32>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
            const double mufu_per_smsp = (double)iters * 8 * (warps / 4);
            printf("mode %d  warps/SM %2d: %.0f cycles, %.2f cycles per warp-MUFU per scheduler (%.2f lanes/clk/SM)\n", mode, warps, c,
                   c / mufu_per_smsp, 4 * 32.0 / (c / mufu_per_smsp));
        }
    return 0;
}
