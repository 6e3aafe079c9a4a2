"""Per-kernel timings on one B200 (CUDA events on the launch stream, warm-up first, inputs cycled so they exceed L2
where the kernel is an HBM streamer).  Development aid; bench.py is the contract."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="attn,gemm_enc,gemm_dec")
    args = ap.parse_args()
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    what = args.what.split(",")

    if "attn" in what:
        B, H, T = 32, 12, 1500
        q = (torch.randn(B, H, T, 64, device="cuda") * 0.3).to(torch.bfloat16)
        k = torch.randn(B, H, T, 64, device="cuda").to(torch.bfloat16)
        v = torch.randn(B, H, T, 64, device="cuda").to(torch.bfloat16)
        out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
        flops = 4.0 * T * T * 64 * B * H
        for tc, name in ((1, "tcgen05"),):
            us = timeit(lambda: _lib.check(lib.wipa_test_enc_attention_bf16(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                                                            B, H, T, tc, st), "attn"), reps=10)
            print(f"enc_attention {name:10s} B={B} H={H} T={T}: {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")

    if "gemm_enc" in what:
        for (M, N, K) in ((48000, 768, 768), (48000, 2304, 768), (48000, 3072, 768), (48000, 768, 3072), (48000, 18432, 768)):
            A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
            Cc = torch.empty(M, N, device="cuda")
            for bn in (128, 0):
                us = timeit(lambda: _lib.check(lib.wipa_test_gemm_bf16(A.data_ptr(), W.data_ptr(), None, Cc.data_ptr(), M, N, K, bn, st), "g"), reps=5)
                print(f"gemm_bf16 M={M} N={N} K={K} bn={bn} (f32 out): {us:9.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
            del A, W, Cc

    if "gemm_dec" in what:
        for M in (16, 64, 256):
            for (N, K) in ((768, 768), (2304, 768), (3072, 768), (768, 3072), (51865, 768)):
                A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
                nW = max(2, int(300e6 / (N * K * 2)))       # cycle through > 2x L2 worth of weights
                Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(nW)]
                Cc = torch.empty(M, N, device="cuda")
                for bn in (32, 64, 128):
                    it = [0]

                    def f():
                        W = Ws[it[0] % nW]
                        it[0] += 1
                        _lib.check(lib.wipa_test_gemm_bf16(A.data_ptr(), W.data_ptr(), None, Cc.data_ptr(), M, N, K, bn, st), "g")
                    us = timeit(f, reps=max(20, nW))
                    print(f"gemm_dec M={M} N={N} K={K} bn={bn}: {us:8.2f} us  {N * K * 2 / us / 1e3:7.1f} GB/s of weights")
                del Ws, A, Cc


if __name__ == "__main__":
    main()
