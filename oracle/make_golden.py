"""Generates tests/golden/* from the real parity oracle (HF transformers Whisper on CPU) and from the reference's own
data / self-test cases.  TEST INFRASTRUCTURE ONLY.  Run in the build container (it reads /root/reference for the IPA
strings; the GPU box never needs /root/reference because the vectors are committed):

    python -m oracle.make_golden

Fixtures
  logmel_{80,128}.npz   HF WhisperFeatureExtractor output for synthetic clips 0..1, frames subsampled 1:25
  tiny_fp32.npz         HF whisper-tiny (seed 0): encoder output (rows 1:50), step-0 logits slice, greedy ids 4 clips x 24
  tiny_gain_fp32.npz    same with init_gain 3.0 (audio-sensitive model): greedy ids + encoder rows
  per_cases.json        the 9 tokenize_ipa assertions (ref:scripts/evaluate_ipa.py:449-457), the 9 printed (ref, hyp)
                        pairs (:387-398) with the restated PER, and 64 reference IPA strings from
                        ref:data/v3_improved/combined_test_ipa.json with seeded-edit hypotheses and their distances
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import hf_reference as hf
from . import per_oracle as po
from . import whisper_oracle as wo

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TOKENIZE_CASES = [["n̩æp", ["n̩", "æ", "p"]], ["ɾ̃æ", ["ɾ̃", "æ"]], ["ə̥tʃ", ["ə̥", "t", "ʃ"]], ["tʃ", ["t", "ʃ"]],
                  ["ŋ̍", ["ŋ̍"]], ["kæt", ["k", "æ", "t"]], ["m̩", ["m̩"]], ["l̩", ["l̩"]], ["", []]]


def printed_pairs():
    """The (ref, hyp) strings of the reference's print-only self test, read from the reference file itself."""
    import ast
    src = open("/root/reference/scripts/evaluate_ipa.py", encoding="utf-8").read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(getattr(t, "id", "") == "test_cases" for t in node.targets):
            return [[e.elts[1].value, e.elts[2].value] for e in node.value.elts]
    raise RuntimeError("test_cases not found in the reference self test")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    audio = wo.synthetic_audio(2)
    for n_mels in (80, 128):
        mel = hf.hf_log_mel(audio, n_mels).numpy()
        np.savez_compressed(os.path.join(GOLD, f"logmel_{n_mels}.npz"), mel_sub=mel[:, :, ::25].astype(np.float32),
                            mel_max=mel.reshape(2, -1).max(1), mel_sum=mel.astype(np.float64).reshape(2, -1).sum(1))
    for name, gain in (("tiny_fp32", 1.0), ("tiny_gain_fp32", 3.0)):
        model = hf.build_hf_model("tiny", seed=0, init_gain=gain)
        audio4 = wo.synthetic_audio(4)
        feats = hf.hf_log_mel(audio4, 80)
        with torch.no_grad():
            enc = model.model.encoder(feats).last_hidden_state
            prompt = torch.tensor([wo.PROMPT_PRE_V3] * 4)
            logits0 = model(input_features=feats, decoder_input_ids=prompt).logits[:, -1].float()
        ids = hf.hf_generate(model, feats, "tiny", max_new=24)
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), enc_rows=enc[:, ::50, :].numpy().astype(np.float32),
                            logits0_slice=logits0[:, ::97].numpy().astype(np.float32),
                            logits0_argmax=logits0.argmax(-1).numpy(), greedy_ids=ids.numpy().astype(np.int64))
    # ---- PER ------------------------------------------------------------------------------------------
    pairs = printed_pairs()
    printed = [[r, h, po.phone_error_rate(r, h)] for r, h in pairs]
    data = json.load(open("/root/reference/data/v3_improved/combined_test_ipa.json", encoding="utf-8"))
    rng = np.random.default_rng(2024)
    picks = rng.choice(len(data), size=64, replace=False)
    corpus = []
    for i in picks:
        ref = data[int(i)]["ipa_transcription"]
        phones = po.tokenize_ipa_fallback(ref)
        hyp = list(phones)
        for _ in range(int(rng.integers(0, 8))):          # seeded substitutions / deletions / insertions
            op = int(rng.integers(0, 3))
            if op == 0 and hyp:
                hyp[int(rng.integers(0, len(hyp)))] = phones[int(rng.integers(0, len(phones)))]
            elif op == 1 and hyp:
                del hyp[int(rng.integers(0, len(hyp)))]
            else:
                hyp.insert(int(rng.integers(0, len(hyp) + 1)), phones[int(rng.integers(0, len(phones)))])
        hyp_s = "".join(hyp)
        hp = po.tokenize_ipa_fallback(hyp_s)
        corpus.append({"ref": ref, "hyp": hyp_s, "n_ref": len(phones), "n_hyp": len(hp),
                       "dist": po.levenshtein_py(phones, hp), "per": po.phone_error_rate(ref, hyp_s)})
    json.dump({"tokenize": TOKENIZE_CASES, "printed_pairs": printed, "corpus": corpus},
              open(os.path.join(GOLD, "per_cases.json"), "w", encoding="utf-8"), ensure_ascii=False, indent=1)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
