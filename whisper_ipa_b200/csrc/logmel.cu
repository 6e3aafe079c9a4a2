// Log-mel front end: reflect-pad + Hann-400 framed real DFT (hop 160) + power + slaney mel filterbank
// + log10/clamp, then the per-clip (max - 8) floor and the (x + 4) / 4 affine.
//
// Replaces mlx_whisper.audio.log_mel_spectrogram (ref:scripts/evaluate_model.py:189,
// ref:scripts/transcribe_single.py:45); arithmetic follows the oracle the task names,
// HF:models/whisper/feature_extraction_whisper.py:135-164 and HF:audio_utils.py:453-544.
//
// Kernel 1 (logmel_stft_mel_kernel): one CTA per (clip, 64-frame tile).  The 10 480 samples the tile needs are
// staged once in shared memory (reflect padding resolved at load).  The windowed DFT is a small dense GEMM
// [402 x 400] x [400 x 64] against a precomputed (window * cos | window * -sin) matrix stored n-major so a
// warp reads 32 consecutive components per load; every warp owns 8 frames and all 402 components
// (13 x 8 register tile), frame samples come from shared memory as warp-wide broadcasts.  The power
// spectrum then overwrites the staging buffer and the filterbank is applied as a banded product
// (each slaney triangle touches <= 32 consecutive bins).  The clip maximum is folded in with one atomic per CTA.
// Kernel 2 (logmel_finish_kernel): elementwise floor + affine.
#include <math.h>
#include <string.h>
#include <vector>

#include "logmel.cuh"

#define LM_TILE_F 64
#define LM_THREADS 256
#define LM_NCOMP 402            // 201 cos + 201 sin
#define LM_NCOMP_PAD 416        // 13 * 32
#define LM_SEG ((LM_TILE_F - 1) * WIPA_HOP + WIPA_N_FFT)     // 10480 samples
#define LM_PW_STRIDE (LM_TILE_F + 1)

static double hz_to_mel(double f) {
    const double logstep = 27.0 / log(6.4);
    return f >= 1000.0 ? 15.0 + log(f / 1000.0) * logstep : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {
    const double logstep = log(6.4) / 27.0;
    return m >= 15.0 ? 1000.0 * exp(logstep * (m - 15.0)) : 200.0 * m / 3.0;
}

int logmel_tables_create(int n_mels, LogmelTables* t) {
    memset(t, 0, sizeof(*t));
    t->n_mels = n_mels;
    // windowed DFT matrix, n-major: dft[n][c], c < 201: w[n] cos(2 pi c n / 400); c >= 201: -w[n] sin(2 pi (c-201) n / 400)
    std::vector<float> dft((size_t)WIPA_N_FFT * LM_NCOMP_PAD, 0.f);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < WIPA_N_FFT; ++n) {
        // torch.hann_window(400) is periodic and evaluated in fp32; reproduce its rounding of the window itself
        const float w = (float)(0.5 - 0.5 * cos(two_pi * n / WIPA_N_FFT));
        for (int k = 0; k < WIPA_N_FREQ; ++k) {
            const int kn = (k * n) % WIPA_N_FFT;
            const double ang = two_pi * kn / WIPA_N_FFT;
            dft[(size_t)n * LM_NCOMP_PAD + k] = (float)((double)w * cos(ang));
            dft[(size_t)n * LM_NCOMP_PAD + WIPA_N_FREQ + k] = (float)(-(double)w * sin(ang));
        }
    }
    // slaney filterbank, float64 like HF then cast to fp32; stored banded
    std::vector<double> pts(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<int> k0(n_mels), klen(n_mels), off(n_mels);
    std::vector<float> wts;
    for (int j = 0; j < n_mels; ++j) {
        const double enorm = 2.0 / (pts[j + 2] - pts[j]);
        int first = -1, last = -1;
        std::vector<float> row(WIPA_N_FREQ);
        for (int k = 0; k < WIPA_N_FREQ; ++k) {
            const double f = 8000.0 * k / (WIPA_N_FREQ - 1);
            const double down = (f - pts[j]) / (pts[j + 1] - pts[j]);
            const double up = (pts[j + 2] - f) / (pts[j + 2] - pts[j + 1]);
            double v = down < up ? down : up;
            v = v > 0.0 ? v : 0.0;
            row[k] = (float)(v * enorm);
            if (row[k] != 0.f) { if (first < 0) first = k; last = k; }
        }
        if (first < 0) { first = 0; last = -1; }
        k0[j] = first; klen[j] = last - first + 1; off[j] = (int)wts.size();
        for (int k = first; k <= last; ++k) wts.push_back(row[k]);
    }
    if (wts.empty()) wts.push_back(0.f);
    std::vector<int> meta(3 * n_mels);
    for (int j = 0; j < n_mels; ++j) { meta[3 * j] = k0[j]; meta[3 * j + 1] = klen[j]; meta[3 * j + 2] = off[j]; }

    WIPA_CUDA_CHECK(cudaMalloc(&t->dft, dft.size() * sizeof(float)));
    WIPA_CUDA_CHECK(cudaMalloc(&t->fb_w, wts.size() * sizeof(float)));
    WIPA_CUDA_CHECK(cudaMalloc(&t->fb_meta, meta.size() * sizeof(int)));
    WIPA_CUDA_CHECK(cudaMemcpy(t->dft, dft.data(), dft.size() * sizeof(float), cudaMemcpyHostToDevice));
    WIPA_CUDA_CHECK(cudaMemcpy(t->fb_w, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice));
    WIPA_CUDA_CHECK(cudaMemcpy(t->fb_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
    return WIPA_OK;
}

void logmel_tables_destroy(LogmelTables* t) {
    if (t->dft) cudaFree(t->dft);
    if (t->fb_w) cudaFree(t->fb_w);
    if (t->fb_meta) cudaFree(t->fb_meta);
    memset(t, 0, sizeof(*t));
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void logmel_init_max_kernel(float* clipmax, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) clipmax[i] = -INFINITY;
}

__global__ void __launch_bounds__(LM_THREADS)
logmel_stft_mel_kernel(const float* __restrict__ audio, const float* __restrict__ dft,
                       const float* __restrict__ fb_w, const int* __restrict__ fb_meta, int n_mels,
                       float* __restrict__ raw, float* __restrict__ clipmax) {
    extern __shared__ float lm_smem[];
    float* seg = lm_smem;                 // LM_SEG floats, later reused as power[201][65]
    __shared__ float red[LM_THREADS / 32];

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * LM_TILE_F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* clip = audio + (size_t)b * WIPA_N_SAMPLES;

    // stage the tile's samples, resolving torch.stft(center=True, pad_mode="reflect")
    const int g0 = f0 * WIPA_HOP - WIPA_N_FFT / 2;
    for (int i = tid; i < LM_SEG; i += LM_THREADS) {
        int g = g0 + i;
        if (g < 0) g = -g;
        if (g >= WIPA_N_SAMPLES) g = 2 * (WIPA_N_SAMPLES - 1) - g;
        g = g < 0 ? 0 : g;     // only reachable for frames >= 3000 of the last tile, which are never stored
        seg[i] = clip[g];
    }
    __syncthreads();

    // DFT as a [402 x 400] x [400 x 8] product per warp, at half the multiply-adds: the periodic Hann window is symmetric
    // (w[n] = w[400 - n], w[0] = 0) and so are the twiddles, dft[400 - n][c] = +dft[n][c] for the cosine components and
    // -dft[n][c] for the sine components, hence
    //     X[c] = dft[200][c] x[200] + sum_{n = 1 .. 199} dft[n][c] (x[n] +/- x[400 - n])
    float acc[13][8];
#pragma unroll
    for (int r = 0; r < 13; ++r)
#pragma unroll
        for (int f = 0; f < 8; ++f) acc[r][f] = 0.f;
    const float* xs = seg + (warp * 8) * WIPA_HOP;
    const float* wcol = dft + lane;
    // component c = lane + 32 r: cosine for c < 201 (r <= 5, and r == 6 for lanes 0 .. 8), sine above
    const bool r6_cos = lane + 192 < WIPA_N_FREQ;
    {
        const float* wr = wcol + (size_t)(WIPA_N_FFT / 2) * LM_NCOMP_PAD;      // n = 200: cos(pi k) = +/-1, the sine row is zero
        float x[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) x[f] = xs[f * WIPA_HOP + WIPA_N_FFT / 2];
#pragma unroll
        for (int r = 0; r < 13; ++r) {
            const float w = __ldg(wr + 32 * r);
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[r][f] = w * x[f];
        }
    }
#pragma unroll 2
    for (int n = 1; n < WIPA_N_FFT / 2; ++n) {
        float xa[8], xd[8], x6[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            const float a = xs[f * WIPA_HOP + n], b = xs[f * WIPA_HOP + WIPA_N_FFT - n];
            xa[f] = a + b;
            xd[f] = a - b;
            x6[f] = r6_cos ? xa[f] : xd[f];
        }
        const float* wr = wcol + (size_t)n * LM_NCOMP_PAD;
#pragma unroll
        for (int r = 0; r < 13; ++r) {
            const float w = __ldg(wr + 32 * r);
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[r][f] = fmaf(w, r < 6 ? xa[f] : (r == 6 ? x6[f] : xd[f]), acc[r][f]);
        }
    }
    __syncthreads();                       // everyone is done reading seg

    // power spectrum into shared memory: first the cos components, then add the sin components
    float* pw = seg;
#pragma unroll
    for (int r = 0; r < 13; ++r) {
        const int c = lane + 32 * r;
        if (c < WIPA_N_FREQ) {
#pragma unroll
            for (int f = 0; f < 8; ++f) pw[c * LM_PW_STRIDE + warp * 8 + f] = acc[r][f] * acc[r][f];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 13; ++r) {
        const int c = lane + 32 * r;
        if (c >= WIPA_N_FREQ && c < LM_NCOMP) {
            const int k = c - WIPA_N_FREQ;
#pragma unroll
            for (int f = 0; f < 8; ++f) pw[k * LM_PW_STRIDE + warp * 8 + f] += acc[r][f] * acc[r][f];
        }
    }
    __syncthreads();

    // banded mel filterbank + log10
    const int f = tid & (LM_TILE_F - 1);
    const int frame = f0 + f;
    float local_max = -INFINITY;
    for (int j = tid / LM_TILE_F; j < n_mels; j += LM_THREADS / LM_TILE_F) {
        const int k0 = fb_meta[3 * j], kl = fb_meta[3 * j + 1], off = fb_meta[3 * j + 2];
        float s = 0.f;
        for (int k = 0; k < kl; ++k) s = fmaf(__ldg(fb_w + off + k), pw[(k0 + k) * LM_PW_STRIDE + f], s);
        const float v = log10f(fmaxf(s, 1e-10f));
        if (frame < WIPA_N_FRAMES) {
            raw[((size_t)b * n_mels + j) * WIPA_N_FRAMES + frame] = v;
            local_max = fmaxf(local_max, v);
        }
    }
    local_max = warp_max(local_max);
    if (lane == 0) red[warp] = local_max;
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
        for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, red[w]);
        atomic_max_float(clipmax + b, m);
    }
}

__global__ void logmel_finish_kernel(float* __restrict__ mel, const float* __restrict__ clipmax, int per_clip,
                                     long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int b = (int)(i / per_clip);
        const float floor_v = clipmax[b] - 8.0f;
        mel[i] = (fmaxf(mel[i], floor_v) + 4.0f) / 4.0f;
    }
}

int launch_logmel(const LogmelTables& t, const float* audio, int B, float* mel, float* clipmax_scratch,
                  cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)(WIPA_N_FREQ * LM_PW_STRIDE > LM_SEG ? WIPA_N_FREQ * LM_PW_STRIDE : LM_SEG);
    static SmemAttr attr;
    WIPA_TRY(wipa_ensure_smem(logmel_stft_mel_kernel, smem, attr));
    logmel_init_max_kernel<<<cdiv(B, 256), 256, 0, st>>>(clipmax_scratch, B);
    WIPA_LAUNCHED();
    dim3 grid(cdiv(WIPA_N_FRAMES, LM_TILE_F), B);
    logmel_stft_mel_kernel<<<grid, LM_THREADS, smem, st>>>(audio, t.dft, t.fb_w, t.fb_meta, t.n_mels, mel, clipmax_scratch);
    WIPA_LAUNCHED();
    const int per_clip = t.n_mels * WIPA_N_FRAMES;
    const long long total = (long long)B * per_clip;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    logmel_finish_kernel<<<blocks, 256, 0, st>>>(mel, clipmax_scratch, per_clip, total);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
