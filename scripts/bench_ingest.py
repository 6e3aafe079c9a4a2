"""Audio-ingest throughput: the same transcription workload fed (a) from audio already in HBM and
(b) from WAV FILES at 48 kHz stereo / 44.1 kHz mono through whisper_ipa_b200.ingest.AudioIngest (worker pool -> pinned
buffers -> one host->device copy per file -> GPU polyphase resampler), what evaluate_model runs per micro-batch.
Prints audio-seconds per second for both and their ratio.  Files are written to --dir first (not timed)."""
import argparse
import json
import os
import sys
import time
import wave

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import whisper_ipa_b200 as w  # noqa: E402
from whisper_ipa_b200 import pipeline  # noqa: E402
from whisper_ipa_b200.ingest import AudioIngest  # noqa: E402
from bench import random_init_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="small")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--rate", type=int, default=48000)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--max-new", type=int, default=220)
    ap.add_argument("--dir", default="/tmp/wipa_ingest_bench")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    os.makedirs(args.dir, exist_ok=True)
    rng = np.random.default_rng(0)
    n_files = args.batch * args.batches
    base = (rng.standard_normal((30 * args.rate, args.channels)) * 0.1 * 32768).clip(-32768, 32767).astype("<i2")
    paths = []
    t0 = time.perf_counter()
    for i in range(n_files):
        p = os.path.join(args.dir, f"clip_{args.rate}_{args.channels}_{i}.wav")
        if not os.path.exists(p):
            with wave.open(p, "wb") as f:
                f.setnchannels(args.channels); f.setsampwidth(2); f.setframerate(args.rate)
                f.writeframes(np.roll(base, i * 997, axis=0).tobytes())
        paths.append(p)
    t_write = time.perf_counter() - t0
    _, sd = random_init_state_dict(args.arch)
    model = w.WhisperIPA(args.arch, dtype="float16", max_batch=args.batch)
    model.load_state_dict(sd)
    tr = pipeline.Transcriber(model, max_new=args.max_new)
    ing = AudioIngest(device=model.device)

    def from_files():
        n = 0
        for idx, audio, errors in ing.iter_batches(paths, args.batch):
            assert not any(errors)
            ids, lens = tr.transcribe_device(audio)
            n += len(idx)
        torch.cuda.synchronize()
        return n

    # warm-up: page cache, pinned pool, kernels
    from_files()
    audio_dev = ing.load_batch(paths[:args.batch])[0].clone()
    torch.cuda.synchronize()

    def resident():
        for _ in range(args.batches):
            tr.transcribe_device(audio_dev)
        torch.cuda.synchronize()

    resident()
    t0 = time.perf_counter(); resident(); t_res = time.perf_counter() - t0
    t0 = time.perf_counter(); n = from_files(); t_files = time.perf_counter() - t0
    out = {"arch": args.arch, "batch": args.batch, "batches": args.batches, "source": f"{args.rate} Hz x {args.channels} ch PCM16 WAV, 30 s",
           "resident_audio_s_per_s": n_files * 30.0 / t_res, "from_files_audio_s_per_s": n * 30.0 / t_files,
           "ratio": (n * 30.0 / t_files) / (n_files * 30.0 / t_res), "file_bytes_per_batch": args.batch * base.nbytes,
           "host_threads": ing.workers, "write_s": t_write}
    print(json.dumps(out))
    ing.close()
    model.close()


if __name__ == "__main__":
    main()
