"""Pins the PER oracle (oracle/per_oracle.{py,c}) to what the reference tree holds for this path."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import per_oracle as po


@pytest.fixture(scope="module")
def cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, "per_cases.json"), encoding="utf-8"))


def test_reference_tokenization_assertions(cases):
    # ref:scripts/evaluate_ipa.py:449-457 — the only value-asserting tests of the reference on this path
    for text, want in cases["tokenize"]:
        assert po.tokenize_ipa_fallback(text) == want


def test_reference_printed_pairs(cases):
    # ref:scripts/evaluate_ipa.py:387-398 (printed upstream); values probed in SURVEY.md §8c
    want = [0.0, 100 / 3, 100 / 3, 100.0, 100 / 3, 100 / 3, 0.0, 50.0, 100 / 3]
    got = [po.phone_error_rate(r, h) for r, h, _ in cases["printed_pairs"]]
    assert got == pytest.approx(want, abs=1e-12)
    assert got == [p for _, _, p in cases["printed_pairs"]]


def test_self_distance_zero_and_empty_rule(cases):
    # ref:scripts/compute_iaa.py:86-89 and ref:scripts/evaluate_ipa.py:96-97
    for c in cases["corpus"][:16]:
        assert po.phone_error_rate(c["ref"], c["ref"]) == 0.0
    assert po.phone_error_rate("", "") == 0.0
    assert po.phone_error_rate("", "kæt") == 100.0
    assert po.phone_error_rate("kæt", "") == 100.0


def test_corpus_fixture(cases):
    for c in cases["corpus"]:
        r, h = po.tokenize_ipa_fallback(c["ref"]), po.tokenize_ipa_fallback(c["hyp"])
        assert (len(r), len(h)) == (c["n_ref"], c["n_hyp"])
        assert po.levenshtein_py(r, h) == c["dist"]
        assert po.phone_error_rate(c["ref"], c["hyp"]) == c["per"]


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 5), max_size=40), st.lists(st.integers(0, 5), max_size=40))
def test_c_matches_python(a, b):
    assert po.levenshtein(np.asarray(a, np.int32), np.asarray(b, np.int32)) == po.levenshtein_py(a, b)


def test_batch_matches_single():
    rng = np.random.default_rng(0)
    refs = [rng.integers(0, 50, size=rng.integers(0, 200)).astype(np.int32) for _ in range(50)]
    hyps = [rng.integers(0, 50, size=rng.integers(0, 200)).astype(np.int32) for _ in range(50)]
    got = po.levenshtein_batch(refs, hyps)
    assert list(got) == [po.levenshtein(r, h) for r, h in zip(refs, hyps)]


def test_evaluate_batch_semantics():
    out = po.evaluate_batch_per([0.0, 50.0, 100.0])
    assert out["per"] == 50.0 and out["num_samples"] == 3
    assert out["per_std"] == pytest.approx(np.sqrt(5000 / 3))


def test_pfer_oracle_properties():
    """PFER restatement (oracle/pfer_oracle.py): with maximally different features the Hamming variant degenerates to the
    unit-cost Levenshtein distance; identical sequences cost 0; a one-feature substitution costs 1/24; the cosine variant
    copies the diagonal for equal vectors and uses the 0.001 guard for zero vectors."""
    import numpy as np
    from oracle import per_oracle as per
    from oracle import pfer_oracle as po
    rng = np.random.default_rng(3)
    n = 12
    feats = {i: np.full(24, 1 if i % 2 else -1) for i in range(n)}      # parity classes: differ in all 24 features or none
    lookup = lambda p: feats[p]
    for _ in range(20):
        r = (rng.integers(0, n // 2, size=int(rng.integers(0, 15))) * 2).tolist()          # even phones
        h = (rng.integers(0, n // 2, size=int(rng.integers(0, 15))) * 2 + 1).tolist()      # odd phones: never equal, all 24 differ
        assert po.pfer_distance_hamming(r, h, lookup) == float(per.levenshtein(r, h))
        assert po.pfer_distance_hamming(r, r, lookup) == 0.0
    one = {0: np.zeros(24), 1: np.eye(24)[0]}
    assert po.pfer_distance_hamming([0], [1], lambda p: one[p]) == 1 / 24
    assert po.pfer_percent(po.pfer_distance_hamming([], [1], lambda p: one[p]), 0, 1) == 100.0
    # cosine: phones 0 and 2 share a vector -> diagonal copy; zero vector -> denominator guard 0.001 -> penalty 1 - 0/0.001 = 1
    cos = {0: np.ones(24), 1: np.zeros(24), 2: np.ones(24)}
    assert po.pfer_distance_cosine([0, 0], [2, 2], lambda p: cos[p]) == 0.0
    assert po.pfer_distance_cosine([0], [1], lambda p: cos[p]) == 1.0
