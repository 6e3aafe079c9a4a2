"""Encoder flash-attention kernel alone (development aid): one whisper-small micro-batch (32 clips x 12 heads x 1500 frames)
through wipa_test_enc_attention_h16, timed with CUDA events.  WIPA_FA_FORM selects the kernel form."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib  # noqa: E402


def main():
    B, H = 32, 12
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.randn(B, H, T, 64, device="cuda", generator=g) * 0.3).half()
    k = torch.randn(B, H, T, 64, device="cuda", generator=g).half()
    v = torch.randn(B, H, T, 64, device="cuda", generator=g).half()
    out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.float16)
    L = _lib.lib("f16")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(L.wipa_test_enc_attention_h16(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, T, 1, st), "fa")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.wipa_test_enc_attention_h16(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, T, 1, st), "fa")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ctas = ((T + 255) // 256) * B * H
    print(f"form {os.environ.get('WIPA_FA_FORM', 'default')} T={T}: {us:.1f} us per launch; {ctas} CTAs of {(T + 127) // 128} key blocks: "
          f"{us * 148 / ctas:.2f} us per CTA")


if __name__ == "__main__":
    main()
