"""Latent cross-attention at the beam-search shape of BASELINE config 4 (development aid): whisper-medium geometry (16 heads),
64 utterances x 5 beams, one list of sequences (beams = 1) against groups of 5 CTAs per key range (beams = 5)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_ipa_b200 import _lib  # noqa: E402


def main():
    H, T, U, K = 16, 1500, 64, 5          # WIPA_XL_WIDE16=0 selects attn_lat.cu instead of attn_lat_wide.cu
    d, S = 64 * H, U * K
    L = _lib.lib("f16")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    E = torch.randn(U, T, d, device="cuda", generator=g).half()
    Et = torch.zeros(U * L.wipa_test_lat_tiled_elems(H, T), device="cuda", dtype=torch.float16)
    _lib.check(L.wipa_test_lat_tile(E.data_ptr(), U, T, H, Et.data_ptr(), st), "lat_tile")
    Qp = (torch.randn(S, H, d, device="cuda", generator=g) * (1.5 / d ** 0.5)).half()
    utt = (torch.arange(S, device="cuda", dtype=torch.int32) // K).to(torch.int32).contiguous()
    C = torch.empty(S, H, d, device="cuda", dtype=torch.float16)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    for beams in (1, K):
        for _ in range(3):
            _lib.check(L.wipa_test_cross_attn_latent(Qp.data_ptr(), Et.data_ptr(), U, utt.data_ptr(), C.data_ptr(), S, H, T, 2, beams, st), "xl")
        tot = 0.0
        reps = 20
        for _ in range(reps):
            flush.zero_()                                    # E (197 MB) must come from HBM again, like between decoder layers
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.wipa_test_cross_attn_latent(Qp.data_ptr(), Et.data_ptr(), U, utt.data_ptr(), C.data_ptr(), S, H, T, 2, beams, st)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"beams arg {beams}: {tot / reps * 1e3:.1f} us per launch ({S} sequences over {U} utterances)")


if __name__ == "__main__":
    main()
