"""One pass of the hot path (log-mel -> encoder -> greedy decode -> PER) for profiling under ncu: random-init weights,
synthetic audio, short decode so the launch list stays small.  Not a benchmark."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import whisper_ipa_b200 as w  # noqa: E402
from whisper_ipa_b200 import metrics, pipeline  # noqa: E402
from bench import random_init_state_dict, synthetic_references  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="small")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--max-new", type=int, default=20)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--dtype", default="float16")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    _, sd = random_init_state_dict(args.arch)
    model = w.WhisperIPA(args.arch, dtype=args.dtype, max_batch=args.batch)
    model.load_state_dict(sd)
    tr = pipeline.Transcriber(model, max_new=args.max_new)
    g = torch.Generator(device="cuda").manual_seed(1)
    audio = torch.randn(args.batch, 480000, device="cuda", generator=g) * 0.1
    refs = synthetic_references(args.batch)
    rf, ro = metrics._pack(refs)
    rf_d, ro_d = torch.from_numpy(rf).cuda(), torch.from_numpy(ro).cuda()
    for _ in range(args.passes):
        ids, lens = tr.transcribe_device(audio)
        counts = tr.score_device(ids, lens, rf_d, ro_d, int(np.max(np.diff(ro))))
    torch.cuda.synchronize()
    print("ok", counts[:2].tolist())


if __name__ == "__main__":
    main()
