// Audio ingest on the GPU: interleaved PCM16 at any source rate -> mono float32 at 16 kHz, zero-padded / cut to 30 s.
//
// Replaces, for a whole micro-batch in one launch, what the reference does per file on the host
// (ref:scripts/evaluate_model.py:187-188, ref:scripts/transcribe_single.py:43-44, ref:scripts/ipa_data_loader.py:48,80):
//   load_audio   : decode -> s16le -> / 32768 -> mono          (mlx_whisper shells out to ffmpeg for decode + resample)
//   pad_or_trim  : zero-pad / truncate to 480 000 samples
// The resampler is the polyphase FIR of scipy.signal.resample_poly(x, up, down) (the host path of audio.load_audio), which
// the tests compare against sample by sample: with h = up * firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down),
// window = ('kaiser', 5.0)) and scipy's padding bookkeeping folded into one constant,
//     out[n] = sum_i h[c(n) - i * up] * x[i],      c(n) = (n + n_pre_remove) * down - n_pre_pad
// over the inputs i with 0 <= c(n) - i * up < n_taps, i.e. about n_taps / up (55 - 61) products per output sample.
//
// Kernel: one CTA per (block of RS_OUT output samples, clip).  The input window the block needs is staged once through
// shared memory with coalesced loads - channels averaged and converted to float on the way in - then every thread forms
// its output from shared memory and the (L1-resident) taps.  HBM traffic = the PCM once + the output once.
#define WIPA_PDL_CLASS 1
#include "common.cuh"

#define RS_OUT 256
#define RS_THREADS 256

__global__ void __launch_bounds__(RS_THREADS)
resample_pcm16_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ clip_off, const int* __restrict__ clip_frames,
                      int n_ch, int up, int down, const float* __restrict__ taps, int n_taps, int c0, int win_cap,
                      float* __restrict__ out, int out_len) {
    extern __shared__ float rs_win[];
    const int clip = blockIdx.y;
    const int n0 = blockIdx.x * RS_OUT;
    const int n_in = clip_frames[clip];
    const int16_t* src = pcm + clip_off[clip];
    float* dst = out + (size_t)clip * out_len;
    // scipy: n_out = ceil(n_in * up / down); beyond it (and beyond out_len) the clip is zero padding
    const long long n_out_ll = ((long long)n_in * up + down - 1) / down;
    const int n_out = n_out_ll < out_len ? (int)n_out_ll : out_len;
    const int n_hi = min(n0 + RS_OUT, out_len);
    if (n0 >= n_out) {
        for (int n = n0 + threadIdx.x; n < n_hi; n += RS_THREADS) dst[n] = 0.f;
        return;
    }
    // inputs touched by outputs [n0, n0 + RS_OUT): i in [ceil((c(n0) - n_taps + 1) / up), floor(c(n0 + RS_OUT - 1) / up)]
    const long long c_first = (long long)n0 * down + c0;
    const long long c_last = (long long)(n0 + RS_OUT - 1) * down + c0;
    long long i_lo = c_first - (n_taps - 1);
    i_lo = i_lo >= 0 ? (i_lo + up - 1) / up : -((-i_lo) / up);
    const long long i_hi = c_last >= 0 ? c_last / up : -1;
    const int win = (int)(i_hi - i_lo + 1);                  // <= win_cap by construction (host computes the same bound)
    const float inv = 1.0f / (32768.0f * (float)n_ch);
    for (int w = threadIdx.x; w < win && w < win_cap; w += RS_THREADS) {
        const long long i = i_lo + w;
        float v = 0.f;
        if (i >= 0 && i < n_in) {
            int acc = 0;
            for (int ch = 0; ch < n_ch; ++ch) acc += (int)src[i * n_ch + ch];
            v = (float)acc * inv;
        }
        rs_win[w] = v;
    }
    __syncthreads();
    const int n = n0 + threadIdx.x;
    if (n < n_hi) {
        float acc = 0.f;
        if (n < n_out) {
            const long long c = (long long)n * down + c0;
            // i runs over the inputs with 0 <= c - i * up < n_taps
            long long ia = c - (n_taps - 1);
            ia = ia >= 0 ? (ia + up - 1) / up : -((-ia) / up);
            const long long ib = c >= 0 ? c / up : -1;
            int k = (int)(c - ib * up);                     // tap index of the newest input, grows by `up` going back
            for (long long i = ib; i >= ia; --i, k += up) acc = fmaf(__ldg(taps + k), rs_win[(int)(i - i_lo)], acc);
        }
        dst[n] = acc;
    }
}

// pcm: device int16, interleaved frames of n_ch channels; clip b occupies elements [clip_off[b], clip_off[b] + clip_frames[b] * n_ch)
// (clip_off: device int64 [B], clip_frames: device int32 [B]); taps: device f32 [n_taps] = up * firwin(...) (or {1} with
// up = down = 1 for 16 kHz sources); c0 = n_pre_remove * down - n_pre_pad of scipy's upfirdn bookkeeping; out: device f32
// [B, out_len] (out_len = 480000 for Whisper), zero-padded beyond each clip's resampled length.
extern "C" int wipa_resample_pcm16(const int16_t* pcm, const long long* clip_off, const int* clip_frames, int B, int n_ch, int up,
                                   int down, const float* taps, int n_taps, int c0, float* out, int out_len, void* stream) {
    WIPA_CHECK(B >= 0 && n_ch >= 1 && n_ch <= 8 && up >= 1 && down >= 1 && n_taps >= 1 && out_len >= 1, WIPA_EINVAL,
               "wipa_resample_pcm16: bad argument (B %d, channels %d, up %d, down %d, taps %d)", B, n_ch, up, down, n_taps);
    if (B == 0) return WIPA_OK;
    WIPA_CHECK(pcm && clip_off && clip_frames && taps && out, WIPA_EINVAL, "wipa_resample_pcm16: null pointer");
    // window of one block: ((RS_OUT - 1) * down + n_taps - 1) / up + 2 inputs
    const long long win_cap = ((long long)(RS_OUT - 1) * down + n_taps - 1) / up + 2;
    const size_t smem = (size_t)win_cap * sizeof(float);
    WIPA_CHECK(smem <= 200 * 1024, WIPA_EUNSUPPORTED, "wipa_resample_pcm16: ratio %d/%d needs %zu bytes of shared memory", up, down, smem);
    static SmemAttr attr;
    if (smem > 48 * 1024) WIPA_TRY(wipa_ensure_smem(resample_pcm16_kernel, smem, attr));
    const dim3 grid(cdiv(out_len, RS_OUT), B);
    resample_pcm16_kernel<<<grid, RS_THREADS, smem, (cudaStream_t)stream>>>(pcm, clip_off, clip_frames, n_ch, up, down, taps, n_taps, c0,
                                                                           (int)win_cap, out, out_len);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
