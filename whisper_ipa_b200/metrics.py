"""Phone-error-rate scoring behind the reference's metric functions (ref:scripts/evaluate_ipa.py:27-105,346-378).

Host side: IPA segmentation, interning of phones to int32 ids, CSR packing, and the final float64 expression
``(distance / len(ref)) * 100.0`` / ``np.mean`` / ``np.std`` exactly as the reference writes them.
Device side: every edit distance (``editdistance.eval`` at ref:scripts/evaluate_ipa.py:100) is computed by libwipa's
warp-per-pair anti-diagonal Levenshtein kernel (csrc/per.cu) through ``wipa_per_batch``.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_panphon_ft = None
_panphon_tried = False


def _feature_table():
    """panphon's FeatureTable when the package exists (the reference's primary segmenter, ref:scripts/evaluate_ipa.py:19-24)."""
    global _panphon_ft, _panphon_tried
    if not _panphon_tried:
        _panphon_tried = True
        try:
            import panphon  # type: ignore
            _panphon_ft = panphon.FeatureTable()
        except Exception:
            _panphon_ft = None
    return _panphon_ft


def _segment_by_category(text: str) -> List[str]:
    # ref:scripts/evaluate_ipa.py:56-65 — combining marks (category M*) and modifier letters U+02B0..U+02FF
    # attach to the preceding base character
    phones: List[str] = []
    for ch in text:
        cat = unicodedata.category(ch)
        attaches = cat[0] == "M" or (cat == "Lm" and 0x02B0 <= ord(ch) <= 0x02FF)
        if attaches and phones:
            phones[-1] = phones[-1] + ch
        else:
            phones.append(ch)
    return phones


def tokenize_ipa(text: str) -> List[str]:
    """IPA string -> list of phones (ref:scripts/evaluate_ipa.py:27-65)."""
    text = text.replace(" ", "")
    if not text:
        return []
    ft = _feature_table()
    if ft is not None:
        segs = ft.ipa_segs(text)
        if "".join(segs) == text:
            return segs
    return _segment_by_category(text)


def normalize_ipa_for_comparison(text: str) -> str:
    """NFC, drop spaces, Latin g -> IPA ɡ (ref:scripts/evaluate_ipa.py:68-77)."""
    text = unicodedata.normalize("NFC", text)
    return text.replace(" ", "").replace("g", "ɡ")


def _pack(seqs: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    if off[-1] >= 2 ** 31:
        raise ValueError("too many ids for int32 offsets")
    flat = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    for i, s in enumerate(seqs):
        if len(s):
            flat[off[i]:off[i + 1]] = np.asarray(s, dtype=np.int32)
    return flat, off.astype(np.int32)


def edit_distance_counts(refs: Sequence[Sequence[int]], hyps: Sequence[Sequence[int]],
                         device: Optional[torch.device] = None) -> torch.Tensor:
    """Batched unit-cost Levenshtein over id sequences on the GPU.
    Returns a device int32 tensor [N, 2] of (edit distance, reference length) — the pair layout the multi-GPU gather moves."""
    assert len(refs) == len(hyps), "Mismatched lengths"        # ref:scripts/evaluate_ipa.py:357
    n = len(refs)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    if n == 0:
        return out
    rf, ro = _pack(refs)
    hf, ho = _pack(hyps)
    max_ref = int(np.max(np.diff(ro))) if n else 0
    bufs = [torch.from_numpy(a).pin_memory().to(dev, non_blocking=True) for a in (rf, ro, hf, ho)]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().wipa_per_batch(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
                                             n, max_ref, out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "wipa_per_batch")
        torch.cuda.current_stream().synchronize()             # pinned staging buffers die with this frame
    return out


def edit_distance_counts_device(ref_flat: torch.Tensor, ref_off: torch.Tensor, hyp_flat: torch.Tensor, hyp_off: torch.Tensor,
                                max_ref_len: int) -> torch.Tensor:
    """Same kernel for inputs already resident in HBM (int32 device tensors, CSR offsets [N+1])."""
    n = ref_off.numel() - 1
    out = torch.zeros((n, 2), dtype=torch.int32, device=ref_flat.device)
    with torch.cuda.device(ref_flat.device):
        _lib.check(_lib.lib().wipa_per_batch(ref_flat.data_ptr(), ref_off.data_ptr(), hyp_flat.data_ptr(), hyp_off.data_ptr(),
                                             n, int(max_ref_len), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "wipa_per_batch")
    return out


def per_from_counts(dist: int, ref_len: int, hyp_len: int) -> float:
    """ref:scripts/evaluate_ipa.py:96-103 — empty-reference rule, then (distance / len(ref)) * 100.0 in float64."""
    if ref_len == 0:
        return 0.0 if hyp_len == 0 else 100.0
    return (dist / ref_len) * 100.0


def per_scores_ids(refs: Sequence[Sequence[int]], hyps: Sequence[Sequence[int]]) -> List[float]:
    counts = edit_distance_counts(refs, hyps).cpu().numpy()
    return [per_from_counts(int(counts[i, 0]), len(refs[i]), len(hyps[i])) for i in range(len(refs))]


def _intern(ref_phones: Sequence[Sequence[str]], hyp_phones: Sequence[Sequence[str]]):
    table: Dict[str, int] = {}
    def ids(seq):
        return [table.setdefault(p, len(table)) for p in seq]
    return [ids(s) for s in ref_phones], [ids(s) for s in hyp_phones]


def phone_error_rate(reference: str, hypothesis: str) -> float:
    """ref:scripts/evaluate_ipa.py:80-105."""
    return phone_error_rates([reference], [hypothesis])[0]


def phone_error_rates(references: Sequence[str], hypotheses: Sequence[str]) -> List[float]:
    assert len(references) == len(hypotheses), "Mismatched lengths"
    rp = [tokenize_ipa(r) for r in references]
    hp = [tokenize_ipa(h) for h in hypotheses]
    ri, hi = _intern(rp, hp)
    return per_scores_ids(ri, hi)


def summarize(per_scores: Sequence[float]) -> Dict:
    """The PER keys of evaluate_batch's dict (ref:scripts/evaluate_ipa.py:370-378): macro average, population std."""
    return {"per": np.mean(per_scores), "per_std": np.std(per_scores), "num_samples": len(per_scores),
            "per_scores": list(per_scores)}


# ---- PFER (ref:scripts/evaluate_ipa.py:108-337) ----------------------------------------------------------------------------
_feature_override: Optional[Dict[str, Sequence[int]]] = None


def set_feature_table(table: Optional[Dict[str, Sequence[int]]]) -> None:
    """Register phone -> 24 numeric features (panphon's `word_to_vector_list(phone, numeric=True)[0]`) for images without
    panphon; phones missing from the table get the zero vector, as the reference does for unknown phones."""
    global _feature_override, _feature_regex
    _feature_override = table
    _feature_regex = None


_feature_regex = None


def load_feature_table(csv_path: str) -> Dict[str, Sequence[int]]:
    """Register panphon's segment table from its ``data/ipa_all.csv`` (columns: ipa, then the 24 features syl ... hireg as
    '+', '-' or '0'), for images where the panphon package is absent but its data file is at hand.  Lookup then follows
    panphon's ``word_to_vector_list(phone, numeric=True)[0]`` (ref:scripts/evaluate_ipa.py:124-134): the phone string is
    scanned with a longest-match regex over the table's segments and the FIRST segment found supplies the vector; no
    segment -> the zero vector."""
    import csv
    import re
    global _feature_regex
    num = {"+": 1, "-": -1, "0": 0}
    table: Dict[str, Sequence[int]] = {}
    with open(csv_path, newline="", encoding="utf-8") as f:
        rows = csv.reader(f)
        header = next(rows)
        if len(header) != 25 or header[0] != "ipa":
            raise ValueError(f"{csv_path}: expected panphon's ipa_all.csv header (ipa + 24 features), got {header[:3]}...")
        for r in rows:
            if len(r) == 25:
                table[unicodedata.normalize("NFD", r[0])] = [num[v] for v in r[1:]]
    set_feature_table(table)
    _feature_regex = re.compile("|".join(re.escape(k) for k in sorted(table, key=len, reverse=True)))
    return table


def _phone_features(phone: str) -> Optional[np.ndarray]:
    if _feature_override is not None:
        v = _feature_override.get(phone)
        if v is None and _feature_regex is not None:               # panphon's segmenting lookup (load_feature_table)
            m = _feature_regex.search(unicodedata.normalize("NFD", phone))
            v = _feature_override.get(m.group(0)) if m else None
        return np.zeros(24, dtype=np.int8) if v is None else np.asarray(v, dtype=np.int8)
    ft = _feature_table()
    if ft is None:
        return None
    try:                                                              # ref:scripts/evaluate_ipa.py:124-134
        vecs = ft.word_to_vector_list(phone, numeric=True)
        return np.asarray(vecs[0], dtype=np.int8) if len(vecs) > 0 else np.zeros(24, dtype=np.int8)
    except Exception:
        return np.zeros(24, dtype=np.int8)


def pfer_available() -> bool:
    return _feature_override is not None or _feature_table() is not None


def feature_edit_distances(ref_ids: Sequence[Sequence[int]], hyp_ids: Sequence[Sequence[int]], feats: np.ndarray,
                           cosine: bool = False, device: Optional[torch.device] = None) -> torch.Tensor:
    """Batched feature-weighted edit distance on the GPU (csrc/pfer.cu): float64 D[len_ref][len_hyp] per pair.
    `feats` is int8 [n_phones, 24] indexed by the interned ids."""
    assert len(ref_ids) == len(hyp_ids), "Mismatched lengths"
    n = len(ref_ids)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.zeros((n,), dtype=torch.float64, device=dev)
    if n == 0:
        return out
    rf, ro = _pack(ref_ids)
    hf, ho = _pack(hyp_ids)
    max_ref = int(np.max(np.diff(ro)))
    f = np.ascontiguousarray(feats, dtype=np.int8).reshape(-1, 24)
    if f.shape[0] == 0:
        f = np.zeros((1, 24), dtype=np.int8)
    bufs = [torch.from_numpy(a).to(dev) for a in (rf, ro, hf, ho, f)]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().wipa_pfer_batch(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), n,
                                              max_ref, bufs[4].data_ptr(), 1 if cosine else 0, out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "wipa_pfer_batch")
        torch.cuda.current_stream().synchronize()
    return out


def phone_feature_error_rates(references: Sequence[str], hypotheses: Sequence[str], cosine: bool = False) -> List[float]:
    """ref:scripts/evaluate_ipa.py:301-315 (Hamming) / :329-343 (cosine) for a batch of string pairs."""
    assert len(references) == len(hypotheses), "Mismatched lengths"
    if not pfer_available():
        raise RuntimeError("PFER needs panphon's feature table: install panphon or call metrics.set_feature_table()")
    rp = [tokenize_ipa(r) for r in references]
    hp = [tokenize_ipa(h) for h in hypotheses]
    table: Dict[str, int] = {}
    def ids(seq):
        return [table.setdefault(p, len(table)) for p in seq]
    ri, hi = [ids(s) for s in rp], [ids(s) for s in hp]
    feats = np.zeros((max(len(table), 1), 24), dtype=np.int8)
    for phone, i in table.items():
        feats[i] = _phone_features(phone)
    d = feature_edit_distances(ri, hi, feats, cosine=cosine).cpu().numpy()
    out = []
    for i in range(len(rp)):
        if len(rp[i]) == 0:                                            # ref:scripts/evaluate_ipa.py:180-181
            out.append(0.0 if len(hp[i]) == 0 else 100.0)
        else:
            out.append((float(d[i]) / len(rp[i])) * 100.0)
    return out


def phone_feature_error_rate(reference: str, hypothesis: str) -> float:
    return phone_feature_error_rates([reference], [hypothesis])[0]


def phone_feature_error_rate_cosine(reference: str, hypothesis: str) -> float:
    return phone_feature_error_rates([reference], [hypothesis], cosine=True)[0]


def evaluate_batch(references: Sequence[str], hypotheses: Sequence[str]) -> Dict:
    """ref:scripts/evaluate_ipa.py:346-378: PER and (Hamming) PFER per pair, macro mean and population std of both.
    Without a phone feature table (panphon absent and none registered) the PFER keys are NaN."""
    assert len(references) == len(hypotheses), "Mismatched lengths"
    per = phone_error_rates(references, hypotheses)
    out = summarize(per)
    if pfer_available():
        pfer = phone_feature_error_rates(references, hypotheses)
        out.update({"pfer": np.mean(pfer), "pfer_std": np.std(pfer), "pfer_scores": list(pfer)})
    else:
        out.update({"pfer": float("nan"), "pfer_std": float("nan"), "pfer_scores": [float("nan")] * len(per)})
    return out
