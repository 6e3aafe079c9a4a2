"""CPU restatement of the reference's phone-feature error rate (PFER) — TEST INFRASTRUCTURE ONLY.

Follows ref:scripts/evaluate_ipa.py line by line with numpy float64, but takes the phone -> feature-vector lookup as an
argument because panphon (ref:requirements.txt:34, panphon==0.22.0) and its `ipa_all.csv` are absent from this image:
  * feature_distance            ref:scripts/evaluate_ipa.py:136-161  (identical phone strings cost 0, else mismatches / 24)
  * PFERCalculator DP           ref:scripts/evaluate_ipa.py:183-213  (insert / delete 1.0, substitute feature_distance)
  * cosine_distance             ref:scripts/evaluate_ipa.py:229-234  (0.001 guard for a zero denominator)
  * PFERCalculatorCosine DP     ref:scripts/evaluate_ipa.py:254-287  (equal vectors copy the diagonal, else min(...) + penalty)
Parity is pinned only structurally (no golden PFER values exist in the reference); unknown phones map to the zero vector
as in get_phone_features (ref:scripts/evaluate_ipa.py:114-134).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

NUM_FEATURES = 24


def feature_distance(p1, p2, feats: Callable) -> float:
    if p1 == p2:
        return 0.0
    f1, f2 = feats(p1), feats(p2)
    mismatches = np.sum(f1 != f2)
    return mismatches / NUM_FEATURES


def pfer_distance_hamming(ref_phones: Sequence, hyp_phones: Sequence, feats: Callable) -> float:
    """dp[m][n] of ref:scripts/evaluate_ipa.py:183-208 (the caller applies the empty-reference rule and the percentage)."""
    m, n = len(ref_phones), len(hyp_phones)
    dp = np.zeros((m + 1, n + 1))
    for i in range(m + 1):
        dp[i][0] = i
    for j in range(n + 1):
        dp[0][j] = j
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            sub_cost = feature_distance(ref_phones[i - 1], hyp_phones[j - 1], feats)
            dp[i][j] = min(dp[i - 1][j] + 1.0, dp[i][j - 1] + 1.0, dp[i - 1][j - 1] + sub_cost)
    return float(dp[m][n])


def cosine_distance(f1: np.ndarray, f2: np.ndarray) -> float:
    denominator = np.linalg.norm(f1) * np.linalg.norm(f2)
    if denominator == 0:
        denominator = 0.001
    return 1.0 - np.dot(f1, f2) / denominator


def pfer_distance_cosine(ref_phones: Sequence, hyp_phones: Sequence, feats: Callable) -> float:
    """dp[m][n] of ref:scripts/evaluate_ipa.py:254-284."""
    rf = [np.asarray(feats(p), dtype=float) for p in ref_phones]
    hf = [np.asarray(feats(p), dtype=float) for p in hyp_phones]
    m, n = len(ref_phones), len(hyp_phones)
    dp = np.zeros((m + 1, n + 1))
    for i in range(m + 1):
        dp[i][0] = i
    for j in range(n + 1):
        dp[0][j] = j
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            if np.array_equal(rf[i - 1], hf[j - 1]):
                dp[i][j] = dp[i - 1][j - 1]
            else:
                penalty = cosine_distance(rf[i - 1], hf[j - 1])
                dp[i][j] = min(dp[i][j - 1], dp[i - 1][j], dp[i - 1][j - 1]) + penalty
    return float(dp[m][n])


def pfer_percent(distance: float, n_ref: int, n_hyp: int) -> float:
    """ref:scripts/evaluate_ipa.py:180-181,211: empty-reference rule, then (dp[m][n] / len(ref)) * 100.0."""
    if n_ref == 0:
        return 0.0 if n_hyp == 0 else 100.0
    return (distance / n_ref) * 100.0
