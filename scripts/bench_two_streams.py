"""Experiment: two half-batch decodes on two streams / host threads vs one full-batch decode (development aid)."""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import whisper_ipa_b200 as w
from whisper_ipa_b200.audio import log_mel_features
from bench import random_init_state_dict

B, MAX_NEW = int(os.environ.get("B", "256")), 100
torch.cuda.set_device(0)
arch = w.ARCHS["small"]
_, sd = random_init_state_dict("small")
g = torch.Generator(device="cuda").manual_seed(1)
audio = torch.randn(B, 480000, device="cuda", generator=g) * 0.1
mel = log_mel_features(audio, arch.n_mels)
prompt = arch.prompt("en", "transcribe", True)
steps = len(prompt) - 1 + MAX_NEW

def run(models, mels, streams):
    def work(m, x, s):
        with torch.cuda.stream(s):
            m.decode_tokens(prompt, MAX_NEW)
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=a) for a in zip(models, mels, streams)]
        for t in ths: t.start()
        for t in ths: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return dt * 1e6 / steps

full = w.WhisperIPA("small", dtype="bfloat16", max_batch=B); full.load_state_dict(sd)
full.encoder(mel, return_features=False)
print(f"one stream  B={B}: {run([full], [mel], [torch.cuda.Stream()]):.1f} us/step")
full.close(); del full
n = 2
halves = [w.WhisperIPA("small", dtype="bfloat16", max_batch=B // n) for _ in range(n)]
for i, m in enumerate(halves):
    m.load_state_dict(sd)
    m.encoder(mel[i * (B // n):(i + 1) * (B // n)], return_features=False)
torch.cuda.synchronize()
print(f"two streams 2x{B // n}: {run(halves, [None] * n, [torch.cuda.Stream() for _ in range(n)]):.1f} us/step (both halves advance one step)")
print(f"one half alone {B // n}: {run(halves[:1], [None], [torch.cuda.Stream()]):.1f} us/step")
