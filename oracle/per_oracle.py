"""CPU oracle for the PER scorer.  TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restates, with the reference line each piece follows:

* ``tokenize_ipa_fallback``  <- ref:scripts/evaluate_ipa.py:56-65 (the Unicode-category branch; the primary
  branch calls ``panphon==0.22.0`` (ref:requirements.txt:34), absent here.  The fallback alone satisfies all 9
  assertions at ref:scripts/evaluate_ipa.py:449-457; beyond them segmentation parity is unpinned)
* ``levenshtein``            <- the contract of ``editdistance.eval`` at ref:scripts/evaluate_ipa.py:100
* ``phone_error_rate``       <- ref:scripts/evaluate_ipa.py:80-105 (incl. the empty-reference rule :96-97 and the
  exact float64 expression ``(distance / len(ref)) * 100.0`` :103)
* ``evaluate_batch_per``     <- ref:scripts/evaluate_ipa.py:346-378 restricted to the PER keys
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import unicodedata
from typing import Dict, List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle() -> str:
    so = os.path.join(_HERE, "_build", "libwipa_oracle.so")
    src = os.path.join(_HERE, "per_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.wipa_oracle_levenshtein.restype = ctypes.c_int32
    return _LIB


def levenshtein_py(a: Sequence, b: Sequence) -> int:
    """Pure-Python two-row DP (small cases only)."""
    if len(a) == 0:
        return len(b)
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, y in enumerate(b, 1):
            cur[j] = min(prev[j - 1] + (x != y), prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[-1]


def levenshtein(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    p = ctypes.POINTER(ctypes.c_int32)
    return int(_lib().wipa_oracle_levenshtein(a.ctypes.data_as(p), len(a), b.ctypes.data_as(p), len(b)))


def levenshtein_batch(refs: Sequence[np.ndarray], hyps: Sequence[np.ndarray]) -> np.ndarray:
    n = len(refs)
    ro = np.zeros(n + 1, np.int32)
    ho = np.zeros(n + 1, np.int32)
    ro[1:] = np.cumsum([len(r) for r in refs])
    ho[1:] = np.cumsum([len(h) for h in hyps])
    r = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int32) for x in refs]) if n else np.zeros(0, np.int32))
    h = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int32) for x in hyps]) if n else np.zeros(0, np.int32))
    out = np.zeros(n, np.int32)
    p = ctypes.POINTER(ctypes.c_int32)
    _lib().wipa_oracle_levenshtein_batch(r.ctypes.data_as(p), ro.ctypes.data_as(p), h.ctypes.data_as(p),
                                         ho.ctypes.data_as(p), n, out.ctypes.data_as(p))
    return out


def tokenize_ipa_fallback(text: str) -> List[str]:
    text = text.replace(" ", "")
    segs: List[str] = []
    for ch in text:
        cat = unicodedata.category(ch)
        mod = cat.startswith("M") or (cat == "Lm" and "ʰ" <= ch <= "˿")
        if segs and mod:
            segs[-1] += ch
        else:
            segs.append(ch)
    return segs


def per_from_counts(dist: int, n_ref: int, n_hyp: int) -> float:
    if n_ref == 0:
        return 0.0 if n_hyp == 0 else 100.0
    return (dist / n_ref) * 100.0


def phone_error_rate(reference: str, hypothesis: str) -> float:
    r = tokenize_ipa_fallback(reference)
    h = tokenize_ipa_fallback(hypothesis)
    if len(r) == 0:
        return 0.0 if len(h) == 0 else 100.0
    return (levenshtein_py(r, h) / len(r)) * 100.0


def per_ids(ref: np.ndarray, hyp: np.ndarray) -> float:
    return per_from_counts(levenshtein(ref, hyp), len(ref), len(hyp))


def evaluate_batch_per(per_scores: Sequence[float]) -> Dict:
    return {"per": np.mean(per_scores), "per_std": np.std(per_scores), "num_samples": len(per_scores),
            "per_scores": list(per_scores)}
