"""GPU parity of the log-mel front end (vs the HF extractor restatement + golden vectors, tolerance 1e-3 relative to the
feature range) and of the PER kernel (bit-exact vs the C oracle, edge cases, and size-independent properties)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_vs_oracle_and_golden(built_lib, golden_dir, n_mels):
    import whisper_ipa_b200 as w
    from oracle import whisper_oracle as wo
    audio = wo.synthetic_audio(3)
    audio[2, 240000:] = 0.0                       # a zero-padded tail exercises the log clamp and the (max - 8) floor
    mel = w.log_mel_features(audio, n_mels).cpu()
    ref = wo.log_mel_spectrogram(torch.from_numpy(audio), n_mels)
    scale = ref.abs().max().item()
    assert (mel - ref).abs().max().item() < 1e-3 * scale
    g = np.load(os.path.join(golden_dir, f"logmel_{n_mels}.npz"))
    assert np.abs(mel[:2, :, ::25].numpy() - g["mel_sub"]).max() < 1e-3 * scale
    # reference-shaped wrapper: [3000, n_mels] for one clip
    one = w.log_mel_spectrogram(audio[0], n_mels=n_mels)
    assert tuple(one.shape) == (3000, n_mels) and torch.equal(one.cpu(), mel[0].T)


def test_logmel_feature_extractor_face(built_lib):
    import whisper_ipa_b200 as w
    from oracle import hf_reference as hf
    rng = np.random.default_rng(5)
    clips = [rng.standard_normal(16000 * 3).astype(np.float32) * 0.05, rng.standard_normal(480000 + 100).astype(np.float32) * 0.2]
    got = w.FeatureExtractor(80)(clips, sampling_rate=16000).input_features.cpu()
    want = hf.hf_log_mel(np.stack([w.pad_or_trim(c) for c in clips]), 80)
    assert (got - want).abs().max().item() < 1e-3 * want.abs().max().item()


def _rand_pairs(rng, n, max_len, vocab):
    refs = [rng.integers(0, vocab, size=rng.integers(0, max_len + 1)).astype(np.int32) for _ in range(n)]
    hyps = [rng.integers(0, vocab, size=rng.integers(0, max_len + 1)).astype(np.int32) for _ in range(n)]
    return refs, hyps


@pytest.mark.parametrize("max_len,vocab", [(8, 2), (40, 5), (120, 100), (224, 51865), (700, 3)])
def test_per_kernel_bit_exact(built_lib, max_len, vocab):
    from whisper_ipa_b200 import metrics
    from oracle import per_oracle as po
    rng = np.random.default_rng(max_len)
    refs, hyps = _rand_pairs(rng, 300, max_len, vocab)
    refs += [np.zeros(0, np.int32), np.zeros(0, np.int32), np.arange(33, dtype=np.int32), np.arange(64, dtype=np.int32)]
    hyps += [np.zeros(0, np.int32), np.arange(5, dtype=np.int32), np.arange(33, dtype=np.int32), np.arange(32, dtype=np.int32)]
    got = metrics.edit_distance_counts(refs, hyps).cpu().numpy()
    want = po.levenshtein_batch(refs, hyps)
    assert (got[:, 0] == want).all()
    assert (got[:, 1] == np.array([len(r) for r in refs])).all()


def test_per_strings_match_golden(built_lib, golden_dir):
    import whisper_ipa_b200 as w
    cases = json.load(open(os.path.join(golden_dir, "per_cases.json"), encoding="utf-8"))
    got = w.phone_error_rates([c["ref"] for c in cases["corpus"]], [c["hyp"] for c in cases["corpus"]])
    assert got == [c["per"] for c in cases["corpus"]]
    printed = cases["printed_pairs"]
    assert w.phone_error_rates([p[0] for p in printed], [p[1] for p in printed]) == [p[2] for p in printed]
    out = w.evaluate_batch([p[0] for p in printed], [p[1] for p in printed])
    assert out["per"] == np.mean([p[2] for p in printed]) and out["num_samples"] == 9
    assert w.phone_error_rate("", "") == 0.0 and w.phone_error_rate("", "kæt") == 100.0 and w.phone_error_rate("kæt", "") == 100.0


def test_per_properties_at_scale(built_lib):
    """Full-size properties (65 536 pairs, lengths up to 224): d(a,a)=0, symmetry, |la-lb| <= d <= max(la,lb), and a
    single substitution / deletion costs exactly 1."""
    from whisper_ipa_b200 import metrics
    rng = np.random.default_rng(99)
    n = 65536
    a = [rng.integers(0, 50, size=rng.integers(1, 225)).astype(np.int32) for _ in range(n)]
    b = [rng.integers(0, 50, size=rng.integers(0, 225)).astype(np.int32) for _ in range(n)]
    dab = metrics.edit_distance_counts(a, b).cpu().numpy()[:, 0]
    dba = metrics.edit_distance_counts(b, a).cpu().numpy()[:, 0]
    daa = metrics.edit_distance_counts(a, a).cpu().numpy()[:, 0]
    la, lb = np.array([len(x) for x in a]), np.array([len(x) for x in b])
    assert (daa == 0).all() and (dab == dba).all()
    assert (dab >= np.abs(la - lb)).all() and (dab <= np.maximum(la, lb)).all()
    sub = [x.copy() for x in a]
    for x in sub:
        x[len(x) // 2] = 1000
    dele = [np.delete(x, len(x) // 2) for x in a]
    assert (metrics.edit_distance_counts(a, sub).cpu().numpy()[:, 0] == 1).all()
    assert (metrics.edit_distance_counts(a, dele).cpu().numpy()[:, 0] == 1).all()


@pytest.mark.parametrize("cosine", [False, True])
def test_pfer_kernel_bit_exact(cosine):
    """Feature-weighted edit distance (PFER): the float64 GPU DP equals the numpy restatement of the reference's DP
    bit for bit, on a synthetic ternary feature table (panphon's table is absent here), incl. empty / long / equal inputs."""
    import numpy as np
    import torch
    from oracle import pfer_oracle as po
    from whisper_ipa_b200 import metrics
    rng = np.random.default_rng(7)
    n_phones = 40
    feats = rng.integers(-1, 2, size=(n_phones, 24)).astype(np.int8)
    feats[3] = feats[5]                          # two distinct phones with identical features
    feats[7] = 0                                 # an "unknown" phone: zero vector (0.001 guard in the cosine variant)
    feats[8] = 0
    lens = [(0, 0), (0, 5), (5, 0), (1, 1), (33, 31), (64, 65), (120, 97), (17, 200)] + \
           [(int(rng.integers(1, 70)), int(rng.integers(1, 70))) for _ in range(40)]
    refs = [rng.integers(0, n_phones, size=a).tolist() for a, _ in lens]
    hyps = [rng.integers(0, n_phones, size=b).tolist() for _, b in lens]
    hyps[4] = list(refs[4][:31])                 # shared prefix
    refs.append(refs[6]); hyps.append(list(refs[6]))     # identical pair -> 0
    got = metrics.feature_edit_distances(refs, hyps, feats, cosine=cosine).cpu().numpy()
    fn = po.pfer_distance_cosine if cosine else po.pfer_distance_hamming
    lookup = lambda p: feats[p]
    want = np.array([fn(r, h, lookup) for r, h in zip(refs, hyps)])
    assert got.dtype == np.float64 and np.array_equal(got, want), f"max diff {np.abs(got - want).max()}"
    assert got[-1] == 0.0


def test_evaluate_batch_with_feature_table():
    """evaluate_batch fills the PFER keys from the GPU scorer once a feature table is registered."""
    import numpy as np
    from oracle import pfer_oracle as po
    from whisper_ipa_b200 import metrics
    table = {"a": [1] * 24, "b": [1] * 12 + [-1] * 12, "t": [-1] * 24, "ʃ": [0] * 23 + [1]}
    metrics.set_feature_table(table)
    try:
        refs, hyps = ["abt", "tʃa", "", "ab"], ["abt", "tab", "a", ""]
        out = metrics.evaluate_batch(refs, hyps)
        lookup = lambda p: np.asarray(table.get(p, [0] * 24))
        want = [po.pfer_percent(po.pfer_distance_hamming(metrics.tokenize_ipa(r), metrics.tokenize_ipa(h), lookup),
                                len(metrics.tokenize_ipa(r)), len(metrics.tokenize_ipa(h))) for r, h in zip(refs, hyps)]
        assert out["pfer_scores"] == want and out["pfer"] == np.mean(want) and out["pfer_std"] == np.std(want)
        assert out["pfer_scores"][0] == 0.0 and out["pfer_scores"][2] == 100.0
    finally:
        metrics.set_feature_table(None)
