/* CPU oracle for the phone-error-rate edit distance.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference computes PER as editdistance.eval(ref_phones, hyp_phones) / len(ref_phones) * 100
 * (ref:scripts/evaluate_ipa.py:100-103).  `editdistance==0.8.1` (ref:requirements.txt:6) is a
 * third-party C++ extension that is NOT vendored under /root/reference and not installed in this
 * image; its published contract is the exact unit-cost Levenshtein distance on element equality
 * (Hyyro's bit-parallel algorithm is only its implementation).  This file restates that contract
 * as the textbook two-row dynamic programme over int32 ids.
 *
 * Parity pinning: ref:scripts/compute_iaa.py:86-89 (d(a,a)==0), the empty-reference rule
 * ref:scripts/evaluate_ipa.py:96-97, and the 9 (ref,hyp) cases of ref:scripts/evaluate_ipa.py:387-398
 * (printed, not asserted, upstream) — see tests/test_oracle_per.py.
 */
#include <stdint.h>
#include <stdlib.h>

int32_t wipa_oracle_levenshtein(const int32_t* a, int32_t na, const int32_t* b, int32_t nb) {
    if (na == 0) return nb;
    if (nb == 0) return na;
    int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nb + 1) * 2);
    int32_t* cur = prev + (nb + 1);
    for (int32_t j = 0; j <= nb; ++j) prev[j] = j;
    for (int32_t i = 1; i <= na; ++i) {
        cur[0] = i;
        const int32_t ai = a[i - 1];
        for (int32_t j = 1; j <= nb; ++j) {
            int32_t sub = prev[j - 1] + (ai != b[j - 1]);
            int32_t del = prev[j] + 1;
            int32_t ins = cur[j - 1] + 1;
            int32_t m = sub < del ? sub : del;
            cur[j] = m < ins ? m : ins;
        }
        int32_t* t = prev; prev = cur; cur = t;
    }
    int32_t d = prev[nb];
    free(prev < cur ? prev : cur);
    return d;
}

/* CSR-packed batch: pair i is ref[ref_off[i]:ref_off[i+1]] vs hyp[hyp_off[i]:hyp_off[i+1]]. */
void wipa_oracle_levenshtein_batch(const int32_t* ref, const int32_t* ref_off, const int32_t* hyp,
                                   const int32_t* hyp_off, int32_t n, int32_t* dist) {
    for (int32_t i = 0; i < n; ++i)
        dist[i] = wipa_oracle_levenshtein(ref + ref_off[i], ref_off[i + 1] - ref_off[i],
                                          hyp + hyp_off[i], hyp_off[i + 1] - hyp_off[i]);
}
