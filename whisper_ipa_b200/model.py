"""The model object the reference scripts hold (``model = load_model(...)``; ``model.encoder(mel)``;
``decode(model, audio_features, options)``) and the HF face ``generate(input_features, ...)``.

Replaces, call for call:
  * ``mlx_whisper.load_models.load_model`` + ``model.set_dtype`` + ``model.update``   ref:scripts/evaluate_model.py:33-73
  * ``model.encoder(mel)`` / ``model.embed_audio(mel)``                               ref:scripts/evaluate_model.py:197, ref:scripts/train_whisper_ipa.py:223
  * ``WhisperForConditionalGeneration.generate`` (greedy, short-form)                 HF:models/whisper/generation_whisper.py:383-968
All arithmetic happens in libwipa (CUDA, sm_100a) behind the C ABI of include/wipa.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .archs import ARCHS, WhisperArch, arch_from_name

T_ENC = 1500
N_FRAMES = 3000
MAX_TARGET = 448

_DTYPES = {"float32": _lib.DTYPE_F32, "fp32": _lib.DTYPE_F32, torch.float32: _lib.DTYPE_F32,
           "float16": _lib.DTYPE_F16, "fp16": _lib.DTYPE_F16, "f16": _lib.DTYPE_F16, torch.float16: _lib.DTYPE_F16,
           "bfloat16": _lib.DTYPE_BF16, "bf16": _lib.DTYPE_BF16, torch.bfloat16: _lib.DTYPE_BF16}
_DTYPE_NAMES = {_lib.DTYPE_F32: "float32", _lib.DTYPE_F16: "float16", _lib.DTYPE_BF16: "bfloat16"}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _i32_array(values: Sequence[int]):
    arr = (C.c_int32 * max(len(values), 1))(*[int(v) for v in values])
    return arr


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Whisper's fixed positional table [length, channels], op for op as HF:models/whisper/modeling_whisper.py:55-65."""
    import math
    log_timescale_increment = math.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length).view(-1, 1) * inv_timescales.view(1, -1)
    return torch.cat([scaled_time.sin(), scaled_time.cos()], dim=1)


class WhisperIPA:
    """One Whisper replica on one GPU (weights + workspaces + KV caches live in a ``wipa_ctx``)."""

    def __init__(self, arch: Union[str, WhisperArch], dtype="float32", max_batch: int = 16, max_beams: int = 1,
                 device: Optional[Union[int, str, torch.device]] = None):
        self.arch = arch if isinstance(arch, WhisperArch) else (ARCHS[arch] if arch in ARCHS else arch_from_name(arch))
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be float32, float16 or bfloat16, got {dtype!r}")
        self.dtype_code = _DTYPES[dtype]
        self.dtype = _DTYPE_NAMES[self.dtype_code]
        self._lib = _lib.lib_for_dtype(self.dtype_code)      # float16 / float32: libwipa.so; bfloat16: libwipa_bf16.so
        if not torch.cuda.is_available():
            raise RuntimeError("whisper_ipa_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.max_batch, self.max_beams = int(max_batch), int(max_beams)
        a = self.arch
        self._arch_c = _lib.Arch(a.d_model, a.enc_layers, a.dec_layers, a.heads, a.ffn, a.n_mels, a.vocab, self.dtype_code)
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_ctx_create(C.byref(self._arch_c), self.max_batch, self.max_beams,
                                                  C.byref(self._ctx)), "wipa_ctx_create")
        # The encoder's position table is a constant, not a learned weight, and MLX-format checkpoints do not carry it.
        # Load HF's own construction of it (float32 torch ops on the host, bit-identical to what WhisperEncoder.__init__
        # writes); a state dict that has "model.encoder.embed_positions.weight" simply overwrites it.  (libwipa pre-fills
        # the slot with double-precision sinusoids for callers of the bare C ABI.)
        self._n_encoded = 0
        self._features_key = None
        self.load_state_dict({"model.encoder.embed_positions.weight": sinusoids(T_ENC, a.d_model)})
        self.suppress_tokens: List[int] = []
        self.begin_suppress_tokens: List[int] = [220, a.eot]

    def _check(self, rc: int, what: str) -> None:
        _lib.check(rc, what, self._lib)

    # ---- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.wipa_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Mapping[str, Union[torch.Tensor, np.ndarray]], strict: bool = False) -> List[str]:
        """Load HF-named tensors (``WhisperForConditionalGeneration.state_dict()`` naming).  May be called again with a
        subset (the reference overlays only ``decoder.*`` keys, ref:scripts/evaluate_model.py:58-73).
        Returns the names that were not recognised (raises on them if ``strict``)."""
        lib = self._lib
        names, tensors, unknown = [], [], []
        for k, v in state_dict.items():
            t = torch.as_tensor(v) if not isinstance(v, torch.Tensor) else v
            names.append(k)
            tensors.append(t.detach().to(device=self.device, dtype=torch.float32).contiguous())
        with torch.cuda.device(self.device):
            for k, t in zip(names, tensors):
                desc = _lib.TensorDesc(k.encode(), t.data_ptr(), t.numel())
                rc = lib.wipa_ctx_load_weights(self._ctx, C.byref(desc), 1, _stream())
                if rc == -1 and b"unknown tensor name" in lib.wipa_last_error():
                    unknown.append(k)
                    continue
                self._check(rc, f"wipa_ctx_load_weights({k})")
            torch.cuda.current_stream().synchronize()          # staging tensors die with this frame
        if strict and unknown:
            raise KeyError(f"unrecognised tensors: {unknown[:5]}{'...' if len(unknown) > 5 else ''}")
        self._n_encoded = 0
        self._features_key = None
        return unknown

    def update(self, params: Mapping[str, Union[torch.Tensor, np.ndarray]]) -> None:
        """Reference-style overlay (``model.update(tree_unflatten(...))``): accepts HF names or MLX names."""
        from .checkpoint import to_hf_state_dict
        self.load_state_dict(to_hf_state_dict(params, self.arch))

    def set_dtype(self, dtype) -> "WhisperIPA":
        if _DTYPES.get(dtype, None) != self.dtype_code:
            raise ValueError("the compute dtype is fixed at construction (weights are converted on load)")
        return self

    # ---- encoder ----------------------------------------------------------------------------------
    def _features_hf_layout(self, mel: torch.Tensor) -> torch.Tensor:
        """Accept the reference layout [B, 3000, n_mels] (ref:scripts/evaluate_model.py:190) or HF's [B, n_mels, 3000]."""
        if isinstance(mel, np.ndarray):
            mel = torch.from_numpy(mel)
        if mel.dim() == 2:
            mel = mel[None]
        n_mels = self.arch.n_mels
        if mel.dim() != 3:
            raise ValueError(f"mel must be 3-D, got {tuple(mel.shape)}")
        if mel.shape[1] == N_FRAMES and mel.shape[2] == n_mels:
            mel = mel.transpose(1, 2)
        elif not (mel.shape[1] == n_mels and mel.shape[2] == N_FRAMES):
            raise ValueError(f"mel shape {tuple(mel.shape)} matches neither [B,{N_FRAMES},{n_mels}] nor [B,{n_mels},{N_FRAMES}] "
                             f"(this model has n_mels={n_mels}; ref:scripts/evaluate_model.py:304-309 defaults --n-mels to 128)")
        return mel.to(device=self.device, dtype=torch.float32).contiguous()

    def encoder(self, mel, return_features: bool = True) -> Optional[torch.Tensor]:
        """mel -> audio features f32 [B, 1500, d]; also projects and caches the cross-attention K/V of these utterances."""
        feats = self._features_hf_layout(mel)
        B = feats.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        out = torch.empty((B, T_ENC, self.arch.d_model), dtype=torch.float32, device=self.device) if return_features else None
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_encode(self._ctx, feats.data_ptr(), B, out.data_ptr() if out is not None else None,
                                              _stream()), "wipa_encode")
        self._n_encoded = B
        self._features_key = None if out is None else (out.data_ptr(), tuple(out.shape), out._version)
        return out

    embed_audio = encoder
    __call_encoder__ = encoder

    def set_audio_features(self, audio_features: torch.Tensor) -> None:
        af = torch.as_tensor(audio_features)
        if (self._features_key is not None and af.is_cuda and self._n_encoded == (af.shape[0] if af.dim() == 3 else 1)
                and self._features_key == (af.data_ptr(), tuple(af.shape), af._version)):
            return          # exactly the tensor encoder() just returned (the reference's `decode(model, model.encoder(mel), ...)`)
        af = af.to(device=self.device, dtype=torch.float32).contiguous()
        if af.dim() == 2:
            af = af[None]
        if af.shape[1:] != (T_ENC, self.arch.d_model):
            raise ValueError(f"audio_features must be [B,{T_ENC},{self.arch.d_model}], got {tuple(af.shape)}")
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_set_audio_features(self._ctx, af.data_ptr(), af.shape[0], _stream()),
                       "wipa_set_audio_features")
        self._n_encoded = af.shape[0]

    # ---- decoder ----------------------------------------------------------------------------------
    def _opts(self, prompt: Sequence[int], max_new: int, suppress: Sequence[int], begin_suppress: Sequence[int]):
        keep = (_i32_array(prompt), _i32_array(suppress), _i32_array(begin_suppress))
        o = _lib.DecodeOpts(C.cast(keep[0], C.POINTER(C.c_int32)), len(prompt), int(max_new), self.arch.eot,
                            C.cast(keep[1], C.POINTER(C.c_int32)), len(suppress),
                            C.cast(keep[2], C.POINTER(C.c_int32)), len(begin_suppress))
        return o, keep

    def decode_tokens(self, prompt: Sequence[int], max_new: int, num_beams: int = 1, length_penalty: float = 1.0,
                      suppress: Optional[Sequence[int]] = None, begin_suppress: Optional[Sequence[int]] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Decode the utterances last passed to ``encoder``.  Returns device int32 (ids [B, max_new] EOT-padded, lengths [B])."""
        B = self._n_encoded
        if B == 0:
            raise RuntimeError("decode before encoder(): no audio features are cached")
        suppress = self.suppress_tokens if suppress is None else list(suppress)
        begin_suppress = self.begin_suppress_tokens if begin_suppress is None else list(begin_suppress)
        o, keep = self._opts(prompt, max_new, suppress, begin_suppress)
        ids = torch.empty((B, max_new), dtype=torch.int32, device=self.device)
        lens = torch.empty((B,), dtype=torch.int32, device=self.device)
        lib = self._lib
        with torch.cuda.device(self.device):
            if num_beams == 1:
                self._check(lib.wipa_decode_greedy(self._ctx, B, C.byref(o), ids.data_ptr(), lens.data_ptr(), _stream()),
                           "wipa_decode_greedy")
            else:
                self._check(lib.wipa_decode_beam(self._ctx, B, int(num_beams), float(length_penalty), C.byref(o),
                                                ids.data_ptr(), lens.data_ptr(), _stream()), "wipa_decode_beam")
        del keep
        return ids, lens

    def teacher_forced_logits(self, tokens: Union[torch.Tensor, np.ndarray]) -> torch.Tensor:
        """Diagnostics: logits f32 [B, T, V] of the cached utterances for forced decoder tokens int [B, T]."""
        tok = np.ascontiguousarray(torch.as_tensor(tokens).cpu().numpy(), dtype=np.int32)
        B, T = tok.shape
        out = torch.empty((B, T, self.arch.vocab), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_decode_logits(self._ctx, B, tok.ctypes.data_as(C.POINTER(C.c_int32)), T,
                                                     out.data_ptr(), _stream()), "wipa_decode_logits")
        return out

    def decode_begin(self, tokens: Union[torch.Tensor, np.ndarray, Sequence[Sequence[int]]]) -> torch.Tensor:
        """Open a stepwise decode over the cached utterances: consume the forced tokens int [B, T] (rows may differ) and
        return the logits that follow the last one, device f32 [B, V].  Continue with decode_next()."""
        tok = np.ascontiguousarray(torch.as_tensor(tokens).cpu().numpy(), dtype=np.int32)
        if tok.ndim == 1:
            tok = tok[None].repeat(self._n_encoded, axis=0)
        B, T = tok.shape
        out = torch.empty((B, self.arch.vocab), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_decode_begin(self._ctx, B, tok.ctypes.data_as(C.POINTER(C.c_int32)), T, out.data_ptr(),
                                                    _stream()), "wipa_decode_begin")
        return out

    def decode_next(self, tokens: torch.Tensor) -> torch.Tensor:
        """Append one token per row (int [B]; a device tensor stays on the device) and return the next logits f32 [B, V]."""
        tok = torch.as_tensor(tokens).to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty((tok.numel(), self.arch.vocab), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.wipa_decode_next(self._ctx, tok.numel(), tok.data_ptr(), out.data_ptr(), _stream()),
                        "wipa_decode_next")
        return out

    def _generate_with_processors(self, prompt: Sequence[int], max_new: int, processors, suppress, begin_suppress):
        """Greedy decoding with caller-side logits processors (HF ``LogitsProcessor.__call__(input_ids, scores)``,
        HF:generation/utils.py:2762-2768): one wipa_decode_next per token, the scores never leave the GPU."""
        B = self._n_encoded
        ids = torch.tensor([list(prompt)] * B, dtype=torch.int64, device=self.device)
        scores = self.decode_begin(ids)
        done = torch.zeros(B, dtype=torch.bool, device=self.device)
        lens = torch.full((B,), max_new, dtype=torch.int32, device=self.device)
        sup = torch.tensor(list(suppress), dtype=torch.int64, device=self.device)
        bsup = torch.tensor(list(begin_suppress), dtype=torch.int64, device=self.device)
        out = torch.full((B, max_new), self.arch.eot, dtype=torch.int32, device=self.device)
        for i in range(max_new):
            if sup.numel():
                scores[:, sup] = -float("inf")
            if i == 0 and bsup.numel():
                scores[:, bsup] = -float("inf")
            for proc in processors:
                scores = proc(ids, scores)
            tok = scores.argmax(dim=-1)
            tok = torch.where(done, torch.full_like(tok, self.arch.eot), tok)
            newly = (~done) & (tok == self.arch.eot)
            lens = torch.where(newly, torch.full_like(lens, i), lens)
            done |= newly
            out[:, i] = tok.to(torch.int32)
            ids = torch.cat([ids, tok[:, None]], dim=1)
            if bool(done.all()) or i + 1 == max_new:
                break
            scores = self.decode_next(tok)
        return out, lens

    @torch.no_grad()
    def generate(self, input_features, decoder_input_ids=None, max_new_tokens: Optional[int] = None,
                 num_beams: int = 1, length_penalty: float = 1.0, language: Optional[str] = None,
                 task: Optional[str] = None, do_sample: bool = False, early_stopping: bool = False,
                 return_dict_in_generate: bool = False, suppress_tokens: Optional[Sequence[int]] = None,
                 begin_suppress_tokens: Optional[Sequence[int]] = None, logits_processor=None, **unused):
        """HF-shaped greedy / beam generation for one 30 s window: returns int64 [B, L] with the prompt and the EOS
        stripped, right-padded with pad_token_id (= EOT), as HF:models/whisper/generation_whisper.py:936-951,1085-1086.
        ``suppress_tokens`` / ``begin_suppress_tokens`` play the role of the checkpoint's ``generation_config`` lists
        (HF:generation/logits_process.py:1812-1902); they default to the model attributes of the same names (empty /
        [220, EOT], what a random-init ``WhisperConfig`` gives HF).  ``max_new_tokens`` defaults to HF's
        ``max_length`` = 448 target positions minus the prompt."""
        if do_sample:
            raise NotImplementedError("sampling is not part of the reference path (temperature 0)")
        if decoder_input_ids is not None:
            p = torch.as_tensor(decoder_input_ids).cpu()
            if p.dim() == 1:
                p = p[None]
            if not bool((p == p[:1]).all()):
                raise NotImplementedError("per-row prompts are not supported; the reference uses one fixed prompt")
            prompt = [int(x) for x in p[0]]
        else:
            prompt = self.arch.prompt(language or "en", task or "transcribe", True)
        max_new = int(max_new_tokens) if max_new_tokens is not None else MAX_TARGET - len(prompt)
        feats = self._features_hf_layout(input_features)
        outs, lens_all = [], []
        for b0 in range(0, feats.shape[0], self.max_batch):
            self.encoder(feats[b0:b0 + self.max_batch], return_features=False)
            if logits_processor:
                if num_beams != 1:
                    raise NotImplementedError("logits_processor is supported with greedy decoding only")
                ids, lens = self._generate_with_processors(
                    prompt, max_new, list(logits_processor),
                    self.suppress_tokens if suppress_tokens is None else suppress_tokens,
                    self.begin_suppress_tokens if begin_suppress_tokens is None else begin_suppress_tokens)
            else:
                ids, lens = self.decode_tokens(prompt, max_new, num_beams=num_beams, length_penalty=length_penalty,
                                               suppress=suppress_tokens, begin_suppress=begin_suppress_tokens)
            outs.append(ids)
            lens_all.append(lens)
        ids = torch.cat(outs).to(torch.int64)
        lens = torch.cat(lens_all)
        L = int(lens.max().item()) if lens.numel() else 0
        seq = ids[:, :L]
        if return_dict_in_generate:
            full = torch.cat([torch.tensor([prompt] * seq.shape[0], dtype=torch.int64, device=seq.device), seq], dim=1)
            return {"sequences": full, "lengths": lens}
        return seq

    def decode(self, mel, options=None):
        from .decoding import decode
        return decode(self, mel, options)

    # ---- introspection ----------------------------------------------------------------------------
    def info(self) -> Dict[str, int]:
        out = {}
        v = C.c_int64()
        for key, sel in (("workspace_bytes", _lib.INFO_WORKSPACE_BYTES), ("crosskv_bytes", _lib.INFO_CROSSKV_BYTES),
                         ("decode_steps", _lib.INFO_DECODE_STEPS), ("xattn_latent", _lib.INFO_XATTN_LATENT)):
            self._check(self._lib.wipa_ctx_get_info(self._ctx, sel, C.byref(v)), "wipa_ctx_get_info")
            out[key] = int(v.value)
        return out


def load_model(path_or_hf_repo: str, dtype="float32", max_batch: int = 16, max_beams: int = 1,
               state_dict: Optional[Mapping[str, torch.Tensor]] = None, device=None) -> WhisperIPA:
    """``mlx_whisper.load_models.load_model`` stand-in (ref:scripts/evaluate_model.py:34): the architecture comes from the
    model id; weights come from ``state_dict`` or a local directory holding ``model.safetensors`` / ``*.npz``."""
    model = WhisperIPA(arch_from_name(path_or_hf_repo), dtype=dtype, max_batch=max_batch, max_beams=max_beams, device=device)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    else:
        from .checkpoint import load_weights_dir
        model.load_state_dict(load_weights_dir(path_or_hf_repo, model.arch))
    return model
