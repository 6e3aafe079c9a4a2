"""Needs libwipa built with -DWIPA_GEMM_DBG: prints the in-kernel timeline (SM clocks) of decode-shaped GEMM launches."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib
lib = _lib.lib()
fn = lib.wipa_debug_gemm_stamps
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
st = torch.cuda.current_stream().cuda_stream
for (M, N, K, bn) in ((256, 768, 768, 32), (64, 768, 768, 32), (256, 768, 3072, 32), (256, 3072, 768, 64)):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(40)]
    b = torch.randn(N, device="cuda")
    Cc = torch.empty(M, N, device="cuda")
    for W in Ws:
        _lib.check(lib.wipa_test_gemm_bf16(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K, bn, st), "g")
    torch.cuda.synchronize()
    ncta = ((N + bn - 1) // bn) * ((M + 127) // 128)
    buf = np.zeros(ncta * 8, dtype=np.uint64)
    assert fn(buf.ctypes.data, ncta * 8) == 0
    t = buf.reshape(ncta, 8).astype(np.int64)
    rel = t - t[:, :1]
    names = ["entry", "W issued", "pdl_wait done", "first full", "mma issued", "acc ready", "epi done", "exit"]
    print(f"M={M} N={N} K={K} bn={bn} ctas={ncta}: median cycles since CTA entry (clock64 is per-SM)")
    for i, n in enumerate(names):
        print(f"   {n:14s} median {np.median(rel[:, i]):9.0f}   max {rel[:, i].max():9.0f}")
