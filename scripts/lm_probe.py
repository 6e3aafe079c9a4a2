"""Log-mel front end alone (development aid): 512 resident 30 s clips, 80 and 128 mel bins, CUDA events."""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_ipa_b200 as w
from whisper_ipa_b200.audio import log_mel_features
g = torch.Generator(device="cuda").manual_seed(1)
audio = torch.randn(512, 480000, device="cuda", generator=g) * 0.1
for n_mels in (80, 128):
    for _ in range(2): log_mel_features(audio, n_mels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): log_mel_features(audio, n_mels)
    e1.record(); torch.cuda.synchronize()
    print(f"log-mel {n_mels} bins, 512 clips: {e0.elapsed_time(e1)/5:.2f} ms")
