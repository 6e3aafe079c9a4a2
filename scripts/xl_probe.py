"""Latent cross-attention kernel alone at the headline shape (development aid): whisper-small geometry, B sequences = B
utterances, chunk-tiled E, through wipa_test_cross_attn_latent; WIPA_LIBWIPA=<other build> gives a same-box A/B."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_ipa_b200 import _lib
H, T, B = 12, 1500, int(sys.argv[1]) if len(sys.argv) > 1 else 512
d = 64 * H
L = _lib.lib("f16"); st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
E = torch.randn(B, T, d, device="cuda", generator=g).half()
Et = torch.zeros(B * L.wipa_test_lat_tiled_elems(H, T), device="cuda", dtype=torch.float16)
_lib.check(L.wipa_test_lat_tile(E.data_ptr(), B, T, H, Et.data_ptr(), st), "tile")
Qp = (torch.randn(B, H, d, device="cuda", generator=g) * (1.5 / d ** 0.5)).half()
utt = torch.arange(B, device="cuda", dtype=torch.int32)
C = torch.empty(B, H, d, device="cuda", dtype=torch.float16)
def run(n):
    for _ in range(n):
        L.wipa_test_cross_attn_latent(Qp.data_ptr(), Et.data_ptr(), B, utt.data_ptr(), C.data_ptr(), B, H, T, 2, 1, st)
run(5); torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(60); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    print(f"{os.environ.get('WIPA_LIBWIPA','default')[-14:]} B={B}: {us:.1f} us per launch, {(B*T*d*2 + 2*B*H*d*2)/us/1e6:.2f} TB/s")
