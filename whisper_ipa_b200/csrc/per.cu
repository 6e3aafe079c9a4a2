// Batched phone-error-rate edit distance: one warp per (reference, hypothesis) pair, anti-diagonal
// wavefront over 32-column strips of the DP matrix.
//
// Replaces editdistance.eval at ref:scripts/evaluate_ipa.py:100 (unit-cost Levenshtein on element equality).
// Integer work, bit-exact by construction; the percentage and the mean/std stay on the host in float64
// (ref:scripts/evaluate_ipa.py:103,370-374).
//
// DP: D[i][j], i over the reference (rows), j over the hypothesis (columns).  Lane l of the warp owns
// column j = 32*s + l + 1 of strip s and walks down the rows; at wavefront step t it fills row i = t - l.
//   up   = its own value of the previous step,
//   left = lane l-1's value of the previous step (one __shfl_up),
//   diag = the `left` it received one step earlier.
// Lane 0 takes left/diag from the boundary column of the previous strip, kept in shared memory and
// overwritten in place by lane 31 (which trails lane 0 by 31 rows, so there is no hazard).
#include "common.cuh"

__global__ void __launch_bounds__(256)
per_levenshtein_kernel(const int32_t* __restrict__ ref, const int32_t* __restrict__ ref_off,
                       const int32_t* __restrict__ hyp, const int32_t* __restrict__ hyp_off, int n_pairs,
                       int words_per_warp, int max_ref_len, int32_t* __restrict__ dist_len) {
    extern __shared__ int32_t smem_per[];
    const int warps_per_cta = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * warps_per_cta + warp;
    if (pair >= n_pairs) return;

    const int r0 = ref_off[pair], nr = ref_off[pair + 1] - r0;
    const int h0 = hyp_off[pair], nh = hyp_off[pair + 1] - h0;
    int32_t* col = smem_per + (size_t)warp * words_per_warp;     // nr + 1 boundary values
    int32_t* sref = col + (nr + 1);                               // nr reference ids

    int result;
    if (nr > max_ref_len || nr < 0 || nh < 0) {
        result = -1;                                              // the caller's bound was wrong: flag the pair, never overrun
    } else if (nr == 0 || nh == 0) {
        result = nr + nh;
    } else {
        for (int i = lane; i <= nr; i += 32) col[i] = i;          // D[i][0]
        for (int i = lane; i < nr; i += 32) sref[i] = ref[r0 + i];
        __syncwarp();
        const int n_strips = (nh + 31) >> 5;
        int cur = 0;
        for (int s = 0; s < n_strips; ++s) {
            const int j = (s << 5) + lane + 1;                    // my column (1-based)
            const bool col_valid = j <= nh;
            const int32_t my_hyp = col_valid ? hyp[h0 + j - 1] : -1;
            cur = j;                                              // D[0][j]
            int diag = j - 1;                                     // D[0][j-1]
            const int steps = nr + 31;
            for (int t = 1; t <= steps; ++t) {
                int left = __shfl_up_sync(0xffffffffu, cur, 1);
                const int i = t - lane;
                const bool active = (i >= 1) && (i <= nr);
                if (lane == 0 && active) left = col[i];
                if (active) {
                    const int sub = diag + (sref[i - 1] != my_hyp ? 1 : 0);
                    const int best = min(sub, min(cur, left) + 1);
                    diag = left;
                    cur = best;
                    if (lane == 31) col[i] = best;                // boundary for the next strip
                }
            }
            __syncwarp();
        }
        // D[nr][nh] sits in the lane that owns column nh of the last strip
        result = __shfl_sync(0xffffffffu, cur, (nh - 1) & 31);
    }
    if (lane == 0) {
        dist_len[2 * pair + 0] = result;
        dist_len[2 * pair + 1] = nr;
    }
}

extern "C" int wipa_per_batch(const int32_t* ref, const int32_t* ref_off, const int32_t* hyp,
                              const int32_t* hyp_off, int N, int max_ref_len, int32_t* dist_len, void* stream) {
    WIPA_CHECK(N >= 0 && max_ref_len >= 0, WIPA_EINVAL, "wipa_per_batch: negative size");
    if (N == 0) return WIPA_OK;
    WIPA_CHECK(ref_off && hyp_off && dist_len, WIPA_EINVAL, "wipa_per_batch: null pointer");
    const int words = 2 * max_ref_len + 2;
    const size_t per_warp = (size_t)words * sizeof(int32_t);
    WIPA_CHECK(per_warp <= 200 * 1024, WIPA_EUNSUPPORTED, "wipa_per_batch: reference longer than %d ids", 25598);
    int warps = (int)((96 * 1024) / per_warp);
    warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
    const size_t smem = per_warp * warps;
    static SmemAttr attr;
    if (smem > 48 * 1024) WIPA_TRY(wipa_ensure_smem(per_levenshtein_kernel, smem, attr));
    const int grid = cdiv(N, warps);
    per_levenshtein_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(ref, ref_off, hyp, hyp_off, N, words,
                                                                              max_ref_len, dist_len);
    WIPA_LAUNCHED();
    return WIPA_OK;
}
