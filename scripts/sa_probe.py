"""Decoder self-attention kernel alone (development aid): S sequences x 12 heads over a scattered paged KV cache at several
lengths, through wipa_test_self_attn, CUDA events.  The pools are sized > L2 so every launch streams from HBM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    H, PAGE, BT = 12, 16, 28
    L = _lib.lib("f16")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(S, H * 64, device="cuda", generator=g) * 0.3
    out = torch.empty(S, H * 64, device="cuda", dtype=torch.float16)
    layers = 12                                            # distinct pools, like the 12 decoder layers of a step
    for length in (8, 32, 64, 114, 160, 224):
        pps = (length + PAGE - 1) // PAGE
        n_pages = S * pps
        pools = [(torch.randn(n_pages, H, PAGE, 64, device="cuda", generator=g).half(),
                  torch.randn(n_pages, H, PAGE, 64, device="cuda", generator=g).half()) for _ in range(layers)]
        perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(1)).reshape(S, pps)
        bt = torch.zeros(S, BT, dtype=torch.int32)
        bt[:, :pps] = perm.to(torch.int32)
        bt = bt.cuda()
        pos = torch.tensor([length - 1], dtype=torch.int32, device="cuda")

        def step():
            for k, v in pools:
                _lib.check(L.wipa_test_self_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), bt.data_ptr(), BT, pos.data_ptr(), out.data_ptr(),
                                                 S, H, 1, st), "self_attn")
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (reps * layers) * 1e3
        mb = S * H * length * 64 * 2 * 2 / 1e6
        print(f"S={S} length {length:3d}: {us:6.1f} us per launch, {mb:6.1f} MB of K/V -> {mb / us:5.2f} TB/s")
        del pools


if __name__ == "__main__":
    main()
