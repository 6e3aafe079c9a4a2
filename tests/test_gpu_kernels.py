"""GPU parity of the standalone kernels through the C ABI: both GEMM families (incl. the conv-as-GEMM overlapping-row
addressing), encoder attention, and the split-K cross-attention streamer."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module")
def lib(built_lib):
    from whisper_ipa_b200 import _lib
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib


# The 16-bit kernels exist in two builds of the same sources: libwipa.so (IEEE fp16, "f16") and libwipa_bf16.so ("bf16").
# Tolerances below are stated for bf16 (8-bit significand); the fp16 build (11-bit) must meet a quarter of them.
@pytest.fixture(params=["f16", "bf16"])
def h16(request):
    return request.param


def _tol(h16, tol_bf16):
    return tol_bf16 if h16 == "bf16" else tol_bf16 / 4


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (200, 136, 240), (1500, 384, 384), (37, 51865 // 50, 384)])
def test_gemm_f32(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    Cc = torch.empty(M, N, device="cuda")
    lib.check(lib.lib().wipa_test_gemm_f32(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K, _st()), "gemm_f32")
    ref = (A.double() @ W.double().T + b.double()).float()
    assert (Cc - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (64, 768, 768), (300, 1000, 384), (1500, 384, 1536), (256, 51865 // 25, 384)])
def test_gemm_h16_tcgen05(lib, h16, bn, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + bn)
    A = torch.randn(M, K, device="cuda", generator=g).to(lib.torch_h16(h16))
    W = torch.randn(N, K, device="cuda", generator=g).to(lib.torch_h16(h16))
    b = torch.randn(N, device="cuda", generator=g)
    Cc = torch.full((M, N), float("nan"), device="cuda")
    lib.check(lib.lib(h16).wipa_test_gemm_h16(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K, bn, _st()),
              "gemm_bf16")
    ref = (A.double() @ W.double().T + b.double()).float()          # exact products of the bf16 operands
    err = (Cc - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (3000, 768, 768), (1500, 1000, 1536), (40000, 520, 384)])
def test_gemm_h16_persistent(lib, h16, M, N, K):
    """block_n = 0 selects the persistent 128 x 256 kernel (double-buffered TMEM accumulators, staged epilogue)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(lib.torch_h16(h16))
    W = torch.randn(N, K, device="cuda", generator=g).to(lib.torch_h16(h16))
    b = torch.randn(N, device="cuda", generator=g)
    Cc = torch.full((M, N), float("nan"), device="cuda")
    lib.check(lib.lib(h16).wipa_test_gemm_h16(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K, 0, _st()), "gemm_bf16")
    ref = (A.float() @ W.float().T + b)
    err = (Cc - ref).abs().max().item()
    assert err < 5e-4 * max(1.0, ref.abs().max().item()), f"max err {err}"


def _gelu(x):
    return torch.nn.functional.gelu(x)          # exact erf form (HF activation_function="gelu")


@pytest.mark.parametrize("bn", [0, 128])
@pytest.mark.parametrize("rpb,nb,N,K", [(1500, 2, 768, 128), (1500, 2, 1152, 64), (200, 3, 384, 192), (128, 1, 192, 64)])
@pytest.mark.parametrize("mode,out_bf16", [(0, 1), (0, 0), (1, 1), (1, 0), (2, 0), (3, 1), (3, 0)])
def test_gemm_epilogue_variants(lib, h16, bn, rpb, nb, N, K, mode, out_bf16):
    """Every epilogue of the encoder GEMMs (bias / GELU / residual / q|k|v head split) on encoder-shaped problems: ragged
    last row tile per batch (1500 = 11 * 128 + 92), a tile that crosses a batch boundary in the output (200 rows), a last
    n-tile that is half empty (1152 = 4 * 256 + 128, 384, 192).  bn = 0 runs the persistent kernel's specialised epilogues."""
    M = rpb * nb
    g = torch.Generator(device="cuda").manual_seed(rpb + N + K + mode)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(lib.torch_h16(h16))
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.2).to(lib.torch_h16(h16))
    b = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    odt = lib.torch_h16(h16) if out_bf16 else torch.float32
    if mode == 3:
        H = N // 192
        out = torch.full((3, nb, H, rpb, 64), float("nan"), device="cuda", dtype=odt)
    else:
        out = torch.full((M, N), float("nan"), device="cuda", dtype=odt)
    if mode == 2:
        out.copy_(resid)                                    # in place, as the encoder's residual stream is updated
    lib.check(lib.lib(h16).wipa_test_gemm_epilogue(A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr() if mode == 2 else 0,
                                                out.data_ptr(), rpb, nb, N, K, mode, out_bf16, bn, _st()), "gemm_epilogue")
    ref = A.double() @ W.double().T + b.double()
    if mode == 1:
        ref = _gelu(ref)
    elif mode == 2:
        ref = ref + resid.double()
    elif mode == 3:
        ref = ref.reshape(nb, rpb, 3, N // 192, 64).permute(2, 0, 3, 1, 4)
    assert not torch.isnan(out.float()).any(), "unwritten output elements"
    scale = max(1.0, ref.abs().max().item())
    err = (out.double() - ref).abs().max().item()
    # fp32 results: accumulation order only; bf16 results: one rounding of the stored value (2^-9 relative) on top
    tol = (_tol(h16, 6e-3) if out_bf16 else 2e-5) * scale
    assert err < tol, f"max err {err} (tol {tol})"


@pytest.mark.parametrize("is_bf16,bn", [(0, 128), (1, 128), (1, 0)])
@pytest.mark.parametrize("C_in,stride,T_out", [(80, 1, 3000), (128, 1, 3000), (384, 2, 1500)])
def test_gemm_conv_rows(lib, h16, is_bf16, bn, C_in, stride, T_out):
    """conv1d(k=3, p=1, stride) as a GEMM over overlapping rows of a zero-row-padded channels-last signal."""
    B, N, T_in = 2, 128, 3000
    g = torch.Generator(device="cuda").manual_seed(C_in + stride)
    x = torch.randn(B, C_in, T_in, device="cuda", generator=g)
    w = torch.randn(N, C_in, 3, device="cuda", generator=g) * 0.1
    dt = lib.torch_h16(h16) if is_bf16 else torch.float32
    rows = torch.zeros(B, T_in + 2, C_in, device="cuda", dtype=dt)
    rows[:, 1:-1] = x.transpose(1, 2).to(dt)
    wg = w.permute(0, 2, 1).reshape(N, 3 * C_in).contiguous().to(dt)           # k = tap * C + c
    out = torch.full((B * T_out, N), float("nan"), device="cuda")
    lib.check(lib.lib(h16).wipa_test_gemm_rows(rows.data_ptr(), is_bf16, stride * C_in, T_out, (T_in + 2) * C_in, B,
                                            wg.data_ptr(), out.data_ptr(), N, 3 * C_in, bn, _st()), "gemm_rows")
    ref = torch.nn.functional.conv1d(rows[:, 1:-1].transpose(1, 2).double(), wg.double().reshape(N, 3, C_in).permute(0, 2, 1),
                                     stride=stride, padding=1).transpose(1, 2).reshape(B * T_out, N).float()
    err = (out - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("use_bf16,tol", [(0, 2e-5), (1, 2e-2), (2, 2e-2)])      # 2 = tcgen05 flash kernel
@pytest.mark.parametrize("T", [1500, 200, 128, 37])
def test_enc_attention(lib, h16, use_bf16, tol, T):
    B, H = 2, 3
    g = torch.Generator(device="cuda").manual_seed(T)
    q = torch.randn(B, H, T, 64, device="cuda", generator=g) * 0.3
    k = torch.randn(B, H, T, 64, device="cuda", generator=g)
    v = torch.randn(B, H, T, 64, device="cuda", generator=g)
    out = torch.empty(B, T, H * 64, device="cuda")
    lib.check(lib.lib(h16).wipa_test_enc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, T, use_bf16, _st()),
              "enc_attention")
    if use_bf16:
        q, k, v = (t.to(lib.torch_h16(h16)).float() for t in (q, k, v))
    p = torch.softmax(q.double() @ k.double().transpose(-1, -2), -1)
    ref = (p @ v.double()).transpose(1, 2).reshape(B, T, H * 64).float()
    assert (out - ref).abs().max().item() < (_tol(h16, tol) if use_bf16 else tol)


@pytest.mark.parametrize("is_bf16,tol", [(0, 2e-5), (1, 2e-2)])
@pytest.mark.parametrize("length", [1, 2, 16, 17, 100, 224, 448])
def test_self_attention_paged(lib, h16, is_bf16, tol, length):
    """One decode step of self-attention over a paged KV cache whose pages are deliberately scattered."""
    B, H, PAGE = 3, 6, 16
    g = torch.Generator(device="cuda").manual_seed(length)
    dt = lib.torch_h16(h16) if is_bf16 else torch.float32
    q = torch.randn(B, H * 64, device="cuda", generator=g) * 0.3
    k = torch.randn(B, H, length, 64, device="cuda", generator=g).to(dt)
    v = torch.randn(B, H, length, 64, device="cuda", generator=g).to(dt)
    pps = (length + PAGE - 1) // PAGE
    n_pages = B * pps + 5
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(1))[:B * pps].reshape(B, pps)
    kpool = torch.full((n_pages, H, PAGE, 64), float("nan"), device="cuda", dtype=dt)
    vpool = torch.full((n_pages, H, PAGE, 64), float("nan"), device="cuda", dtype=dt)
    for b in range(B):
        for pg in range(pps):
            n = min(PAGE, length - pg * PAGE)
            kpool[perm[b, pg], :, :n] = k[b, :, pg * PAGE:pg * PAGE + n]
            vpool[perm[b, pg], :, :n] = v[b, :, pg * PAGE:pg * PAGE + n]
    bt = torch.zeros(B, 28, dtype=torch.int32)
    bt[:, :pps] = perm.to(torch.int32)
    bt = bt.cuda()
    pos = torch.tensor([length - 1], dtype=torch.int32, device="cuda")
    out = torch.empty(B, H * 64, device="cuda", dtype=dt)
    lib.check(lib.lib(h16).wipa_test_self_attn(q.data_ptr(), kpool.data_ptr(), vpool.data_ptr(), bt.data_ptr(), 28, pos.data_ptr(),
                                            out.data_ptr(), B, H, is_bf16, _st()), "self_attn")
    qh = q.double().view(B, H, 1, 64)
    ref = (torch.softmax(qh @ k.double().transpose(-1, -2), -1) @ v.double()).reshape(B, H * 64)
    err = (out.double() - ref).abs().max().item()
    assert err < (_tol(h16, tol) if is_bf16 else tol), f"max err {err}"


@pytest.mark.parametrize("dtype,tol", [("float32", 2e-5), ("bfloat16", 2e-2), ("float16", 5e-3)])
def test_cross_attention_streamer(lib, tiny_sd, dtype, tol, monkeypatch):
    """Stream-K cross-attention over the per-layer cross-KV cache (the fp32 path, and the bf16 path with WIPA_XATTN_LATENT=0)."""
    import whisper_ipa_b200 as w
    monkeypatch.setenv("WIPA_XATTN_LATENT", "0")
    B = 5
    m = w.WhisperIPA("tiny", dtype=dtype, max_batch=B)
    m.load_state_dict(tiny_sd)
    g = torch.Generator(device="cuda").manual_seed(3)
    enc = torch.randn(B, 1500, 384, device="cuda", generator=g)
    m.set_audio_features(enc)
    H, d = 6, 384
    for layer in (0, 3):
        q = torch.randn(B, d, device="cuda", generator=g)
        out = torch.empty(B, d, device="cuda")
        lib.check(m._lib.wipa_test_cross_attn(m._ctx, B, layer, q.data_ptr(), out.data_ptr(), _st()), "cross_attn")
        lp = f"model.decoder.layers.{layer}.encoder_attn."
        wk, wv, bv = (tiny_sd[lp + n].cuda() for n in ("k_proj.weight", "v_proj.weight", "v_proj.bias"))
        e = enc
        tdt = {"bfloat16": torch.bfloat16, "float16": torch.float16}.get(dtype)
        if tdt is not None:
            e, wk, wv = (t.to(tdt).float() for t in (e, wk, wv))
        K = (e.double() @ wk.double().T).view(B, 1500, H, 64).transpose(1, 2)
        V = (e.double() @ wv.double().T + bv.double()).view(B, 1500, H, 64).transpose(1, 2)
        if tdt is not None:
            K, V = K.float().to(tdt).double(), V.float().to(tdt).double()
        qh = q.double().view(B, H, 1, 64)
        ref = (torch.softmax(qh @ K.transpose(-1, -2), -1) @ V).reshape(B, d).float()
        err = (out - ref).abs().max().item()
        assert err < tol, f"layer {layer}: max err {err}"
    m.close()


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("H,T,S,U", [(12, 1500, 5, 3), (12, 100, 200, 4), (6, 1500, 3, 3), (16, 333, 4, 2), (8, 48, 2, 1), (12, 1500, 300, 150),
                                     (20, 1500, 5, 3), (20, 100, 200, 4), (20, 333, 1, 1), (20, 1500, 160, 40)])
def test_cross_attention_latent(lib, h16, H, T, S, U, layout):
    """`layout`: how E reaches the kernel - 0 row-major behind a tensor map (TMA boxes), 1 / 2 the chunk-tiled, pre-swizzled
    image the context keeps (bulk copies; converted inside the call / beforehand by wipa_test_lat_tile).
    Latent cross-attention kernel (mma.sync over TMA-swizzled tiles of the encoder output): C = softmax(Q' E^T) E per
    sequence, sequences mapped to utterances (beams share E), more sequences than SMs, ragged last key chunk, both key
    chunk sizes (48 keys up to 12 heads, 32 above).  20 heads (whisper-large*) run attn_lat_wide.cu: two CTAs of 10 heads per
    (sequence, chunk) range, chunk-tiled layout only."""
    if H == 20 and layout == 0:
        pytest.skip("20 heads: the kernel reads the chunk-tiled layout only")
    d = 64 * H
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + T + S)
    E = torch.randn(U, T, d, device="cuda", generator=g).to(lib.torch_h16(h16))
    Qp = (torch.randn(S, H, d, device="cuda", generator=g) * (1.5 / d ** 0.5)).to(lib.torch_h16(h16))
    utt = (torch.arange(S, device="cuda", dtype=torch.int32) * U // S).to(torch.int32).contiguous()
    C = torch.full((S, H, d), float("nan"), device="cuda", dtype=lib.torch_h16(h16))
    Ein = E
    if layout == 2:
        Ein = torch.zeros(U * lib.lib(h16).wipa_test_lat_tiled_elems(H, T), device="cuda", dtype=lib.torch_h16(h16))
        lib.check(lib.lib(h16).wipa_test_lat_tile(E.data_ptr(), U, T, H, Ein.data_ptr(), _st()), "lat_tile")
    lib.check(lib.lib(h16).wipa_test_cross_attn_latent(Qp.data_ptr(), Ein.data_ptr(), U, utt.data_ptr(), C.data_ptr(), S, H, T, layout, 1, _st()),
              "cross_attn_latent")
    Eu = E.float()[utt.long()]                                          # [S, T, d]
    scores = torch.einsum("shd,std->sht", Qp.float(), Eu)
    P = torch.softmax(scores, dim=-1)
    ref = torch.einsum("sht,std->shd", P, Eu)
    assert not torch.isnan(C.float()).any()
    err = (C.float() - ref).abs().max().item()
    # bf16 probabilities (2^-9 relative each, averaged over the keys) + one bf16 rounding of the result
    assert err < _tol(h16, 1.5e-2) * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("H,S", [(12, 256), (12, 512), (12, 5), (6, 130), (8, 300), (16, 64), (12, 1000)])
def test_xlq_fused(lib, h16, H, S):
    """Absorbed cross-attention queries in one kernel (gemm_q2.cu): out[s, h] = Wk_h^T (Wq_h a_s + b_h), the q rows rounded
    to 16 bits between the two products exactly like the two-node path (q in h16 through L2).  Ragged last M tile, both
    chunk widths (128 columns up to 256 sequences, 256 above when d allows), 6 / 8 / 12 / 16 heads."""
    d = 64 * H
    dt = lib.torch_h16(h16)
    g = torch.Generator(device="cuda").manual_seed(H * 100 + S)
    A = torch.randn(S, d, device="cuda", generator=g).to(dt)
    Wq = (torch.randn(d, d, device="cuda", generator=g) / d ** 0.5).to(dt)
    Wk = (torch.randn(d, d, device="cuda", generator=g) / d ** 0.5).to(dt)
    b = torch.randn(d, device="cuda", generator=g) * 0.1
    scratch = torch.empty(H * d * 64, device="cuda", dtype=dt)
    out = torch.full((S, H, d), float("nan"), device="cuda", dtype=dt)
    lib.check(lib.lib(h16).wipa_test_xlq_fused(A.data_ptr(), Wq.data_ptr(), Wk.data_ptr(), b.data_ptr(), scratch.data_ptr(),
                                               out.data_ptr(), S, H, _st()), "xlq_fused")
    q = (A.float() @ Wq.float().T + b).to(dt).float().view(S, H, 64)             # the 16-bit q tile
    ref = torch.einsum("shi,hij->shj", q, Wk.float().view(H, 64, d))
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs().max().item()
    # fp32 accumulation over K = d and K = 64; the error is the final 16-bit rounding plus q values that land on the other
    # side of a rounding boundary (accumulation order differs from torch's)
    assert err < _tol(h16, 2e-2) * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("H,T,U,K", [(16, 1500, 64, 5), (12, 1500, 7, 5), (16, 333, 3, 2), (8, 100, 40, 8), (12, 1500, 31, 4)])
def test_cross_attention_latent_beam_groups(lib, h16, H, T, U, K):
    """Beam search: the K beams of an utterance are adjacent sequences with the same E.  The kernel then cuts the list of
    (utterance, chunk) units into ranges walked by GROUPS of K CTAs, one beam per CTA (E leaves HBM once per utterance);
    results per sequence must not depend on that: compared with the same fp32 reference as the greedy layout."""
    d = 64 * H
    S = U * K
    dt = lib.torch_h16(h16)
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + T + S)
    E = torch.randn(U, T, d, device="cuda", generator=g).to(dt)
    Qp = (torch.randn(S, H, d, device="cuda", generator=g) * (1.5 / d ** 0.5)).to(dt)
    utt = (torch.arange(S, device="cuda", dtype=torch.int32) // K).to(torch.int32).contiguous()
    C = torch.full((S, H, d), float("nan"), device="cuda", dtype=dt)
    Et = torch.zeros(U * lib.lib(h16).wipa_test_lat_tiled_elems(H, T), device="cuda", dtype=dt)
    lib.check(lib.lib(h16).wipa_test_lat_tile(E.data_ptr(), U, T, H, Et.data_ptr(), _st()), "lat_tile")
    lib.check(lib.lib(h16).wipa_test_cross_attn_latent(Qp.data_ptr(), Et.data_ptr(), U, utt.data_ptr(), C.data_ptr(), S, H, T, 2, K, _st()),
              "cross_attn_latent")
    Eu = E.float()[utt.long()]
    P = torch.softmax(torch.einsum("shd,std->sht", Qp.float(), Eu), dim=-1)
    ref = torch.einsum("sht,std->shd", P, Eu)
    assert not torch.isnan(C.float()).any()
    err = (C.float() - ref).abs().max().item()
    assert err < _tol(h16, 1.5e-2) * max(1.0, ref.abs().max().item()), f"max err {err}"
