// Batched phone-feature error rate (PFER): feature-weighted edit distance, one warp per (reference, hypothesis) pair,
// the same anti-diagonal wavefront over 32-column strips as per.cu but with float64 cells.
//
// Replaces the pure-Python O(m n) DPs of the reference:
//   mode 0  PFERCalculator.phone_feature_error_rate          (ref:scripts/evaluate_ipa.py:163-213)
//           insert / delete cost 1.0, substitution cost 0.0 for identical phones, else mismatching features / 24
//           (feature_distance, ref:scripts/evaluate_ipa.py:136-161)
//   mode 1  PFERCalculatorCosine.phone_feature_error_rate    (ref:scripts/evaluate_ipa.py:236-287)
//           equal FEATURE VECTORS copy the diagonal; otherwise min(insert, delete, substitute) + (1 - cos_sim), with the
//           reference's 0.001 guard for a zero denominator (cosine_distance, :229-234)
// Every operation is an IEEE float64 add / min / divide / sqrt applied in the reference's order, so D[m][n] is
// bit-identical to the numpy result; the percentage (D / len(ref)) * 100.0 stays on the host like the PER one.
// Phones arrive interned as int32 ids; feats is int8 [n_phones, 24] (panphon's numeric features, all-zero rows for
// unknown phones as in get_phone_features, :114-134).
#define WIPA_PDL_CLASS 1
#include "common.cuh"

#define PFER_NF 24

__device__ __forceinline__ double pfer_cost(int mode, int ref_id, int hyp_id, const int8_t* __restrict__ fr,
                                            const int8_t* __restrict__ fh, bool* same_vec) {
    int mism = 0, dot = 0, nr2 = 0, nh2 = 0;
#pragma unroll
    for (int k = 0; k < PFER_NF; ++k) {
        const int a = fr[k], b = fh[k];
        mism += (a != b);
        dot += a * b;
        nr2 += a * a;
        nh2 += b * b;
    }
    *same_vec = (mism == 0);
    if (mode == 0) return ref_id == hyp_id ? 0.0 : (double)mism / 24.0;
    double den = sqrt((double)nr2) * sqrt((double)nh2);
    if (den == 0.0) den = 0.001;
    return 1.0 - (double)dot / den;
}

__global__ void __launch_bounds__(128)
pfer_kernel(const int32_t* __restrict__ ref, const int32_t* __restrict__ ref_off, const int32_t* __restrict__ hyp,
            const int32_t* __restrict__ hyp_off, int n_pairs, const int8_t* __restrict__ feats, int mode,
            int bytes_per_warp, int max_ref_len, double* __restrict__ dist) {
    extern __shared__ __align__(8) uint8_t smem_pfer[];
    const int warps_per_cta = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * warps_per_cta + warp;
    if (pair >= n_pairs) return;
    const int r0 = ref_off[pair], nr = ref_off[pair + 1] - r0;
    const int h0 = hyp_off[pair], nh = hyp_off[pair + 1] - h0;
    uint8_t* base = smem_pfer + (size_t)warp * bytes_per_warp;
    double* col = reinterpret_cast<double*>(base);                          // nr + 1 boundary values
    int32_t* sref = reinterpret_cast<int32_t*>(col + (nr + 1));             // nr reference ids
    int8_t* sfeat = reinterpret_cast<int8_t*>(sref + nr);                   // nr x 24 reference features

    double result;
    if (nr > max_ref_len || nr < 0 || nh < 0) {
        result = -1.0;                                  // the caller's bound was wrong: flag the pair, never overrun
    } else if (nr == 0 || nh == 0) {
        result = (double)(nr + nh);                     // first row / column of the DP table
    } else {
        for (int i = lane; i <= nr; i += 32) col[i] = (double)i;
        for (int i = lane; i < nr; i += 32) sref[i] = ref[r0 + i];
        __syncwarp();
        for (int i = lane; i < nr * PFER_NF; i += 32) sfeat[i] = feats[(size_t)sref[i / PFER_NF] * PFER_NF + i % PFER_NF];
        __syncwarp();
        const int n_strips = (nh + 31) >> 5;
        double cur = 0.0;
        for (int s = 0; s < n_strips; ++s) {
            const int j = (s << 5) + lane + 1;
            const bool col_valid = j <= nh;
            const int32_t my_hyp = col_valid ? hyp[h0 + j - 1] : 0;
            int8_t fh[PFER_NF];
#pragma unroll
            for (int k = 0; k < PFER_NF; ++k) fh[k] = col_valid ? feats[(size_t)my_hyp * PFER_NF + k] : (int8_t)0;
            cur = (double)j;                            // D[0][j]
            double diag = (double)(j - 1);              // D[0][j-1]
            const int steps = nr + 31;
            for (int t = 1; t <= steps; ++t) {
                double left = __shfl_up_sync(0xffffffffu, cur, 1);
                const int i = t - lane;
                const bool active = (i >= 1) && (i <= nr);
                if (lane == 0 && active) left = col[i];
                if (active) {
                    bool same_vec;
                    const double c = pfer_cost(mode, sref[i - 1], my_hyp, sfeat + (size_t)(i - 1) * PFER_NF, fh, &same_vec);
                    double best;
                    if (mode == 0) {
                        // min(dp[i-1][j] + 1.0, dp[i][j-1] + 1.0, dp[i-1][j-1] + sub_cost)
                        best = fmin(fmin(cur + 1.0, left + 1.0), diag + c);
                    } else if (same_vec) {
                        best = diag;
                    } else {
                        best = fmin(fmin(left, cur), diag) + c;
                    }
                    diag = left;
                    cur = best;
                    if (lane == 31) col[i] = best;
                }
            }
            __syncwarp();
        }
        result = __shfl_sync(0xffffffffu, cur, (nh - 1) & 31);
    }
    if (lane == 0) dist[pair] = result;
}

extern "C" int wipa_pfer_batch(const int32_t* ref, const int32_t* ref_off, const int32_t* hyp, const int32_t* hyp_off, int N,
                               int max_ref_len, const int8_t* feats, int mode, double* dist, void* stream) {
    WIPA_CHECK(N >= 0 && max_ref_len >= 0 && (mode == 0 || mode == 1), WIPA_EINVAL, "wipa_pfer_batch: bad size / mode");
    if (N == 0) return WIPA_OK;
    WIPA_CHECK(ref_off && hyp_off && feats && dist, WIPA_EINVAL, "wipa_pfer_batch: null pointer");
    size_t per_warp = (size_t)(max_ref_len + 1) * 8 + (size_t)max_ref_len * 4 + (size_t)max_ref_len * PFER_NF;
    per_warp = (per_warp + 15) & ~(size_t)15;
    WIPA_CHECK(per_warp <= 200 * 1024, WIPA_EUNSUPPORTED, "wipa_pfer_batch: reference longer than %d phones", 5600);
    int warps = (int)((96 * 1024) / per_warp);
    warps = warps < 1 ? 1 : (warps > 4 ? 4 : warps);
    const size_t smem = per_warp * warps;
    static SmemAttr attr;
    if (smem > 48 * 1024) WIPA_TRY(wipa_ensure_smem(pfer_kernel, smem, attr));
    pfer_kernel<<<cdiv(N, warps), warps * 32, smem, (cudaStream_t)stream>>>(ref, ref_off, hyp, hyp_off, N, feats, mode,
                                                                            (int)per_warp, max_ref_len, dist);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
