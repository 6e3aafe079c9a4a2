"""Decode-step timing alone (development aid): encode B synthetic clips once, then time greedy decodes.
Prints microseconds per decode step.  Env knobs of libwipa (WIPA_BN_DEC, WIPA_CA_SPLIT, WIPA_PDL, ...) apply."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import whisper_ipa_b200 as w  # noqa: E402
from whisper_ipa_b200.audio import log_mel_features  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="small")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--max-new", type=int, default=100)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    arch = w.ARCHS[args.arch]
    # random weights with the right names/shapes without building the HF module (fast)
    from bench import random_init_state_dict
    _, sd = random_init_state_dict(args.arch)
    model = w.WhisperIPA(args.arch, dtype="float16", max_batch=args.batch)
    model.load_state_dict(sd)
    g = torch.Generator(device="cuda").manual_seed(1)
    audio = torch.randn(args.batch, 480000, device="cuda", generator=g) * 0.1
    mel = log_mel_features(audio, arch.n_mels)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.encoder(mel, return_features=False)
    torch.cuda.synchronize()
    e0.record()
    model.encoder(mel, return_features=False)
    e1.record()
    torch.cuda.synchronize()
    t_enc = e0.elapsed_time(e1)
    prompt = arch.prompt("en", "transcribe", True)
    ids, lens = model.decode_tokens(prompt, args.max_new)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        ids, lens = model.decode_tokens(prompt, args.max_new)
    e1.record()
    torch.cuda.synchronize()
    steps = len(prompt) - 1 + args.max_new
    us = 1000.0 * e0.elapsed_time(e1) / args.reps / steps
    print(f"[{args.tag}] B={args.batch} arch={args.arch}: encoder {t_enc:.1f} ms, decode {us:.1f} us/step over {steps} steps "
          f"(ids checksum {int(ids.sum().item())})")


if __name__ == "__main__":
    main()
