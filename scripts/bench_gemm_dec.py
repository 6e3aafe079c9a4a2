"""Decode-GEMM node latency inside a CUDA graph (development aid).  Sweeps tile width and rows-per-tile through
wipa_test_gemm_rows; weights cycle through > L2 bytes; PDL as configured by WIPA_PDL."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib  # noqa: E402


def main():
    lib = _lib.lib()
    torch.cuda.set_device(0)
    s = torch.cuda.Stream()
    for M in (64, 256):
        for (N, K) in ((768, 768), (2304, 768), (3072, 768), (768, 3072)):
            A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            nW = max(4, int(400e6 / (N * K * 2)))
            Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(nW)]
            Cc = torch.empty(M, N, device="cuda")
            for bn in (32, 64, 128):
                for rpb in (64, 128):
                    if rpb > M:
                        continue
                    nb = M // rpb
                    with torch.cuda.stream(s):
                        st = s.cuda_stream

                        def launch_all():
                            for Wt in Ws:
                                _lib.check(lib.wipa_test_gemm_rows(A.data_ptr(), 1, K, rpb, rpb * K, nb, Wt.data_ptr(), Cc.data_ptr(),
                                                                   N, K, bn, st), "g")
                        launch_all()
                        s.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=s):
                            launch_all()
                        g.replay()
                        s.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(s)
                        for _ in range(3):
                            g.replay()
                        e1.record(s)
                        s.synchronize()
                    us = e0.elapsed_time(e1) * 1000 / 3 / nW
                    print(f"M={M:3d} N={N:5d} K={K:4d} bn={bn:3d} rows/tile={rpb:3d}: {us:7.2f} us/node  ({N * K * 2 / us / 1e3:6.0f} GB/s of W)", flush=True)
            del Ws


if __name__ == "__main__":
    main()
