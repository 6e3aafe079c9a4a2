"""Long-form transcription: the ``mlx_whisper.transcribe(audio_path, path_or_hf_repo=..., language=..., word_timestamps=False)``
call of the reference's BASE-MODEL branch (ref:scripts/evaluate_model.py:112-119; ``transcribe_with_model`` :82-124).

mlx-whisper 0.4.3 is not vendored under /root/reference, so this restates its published algorithm (the OpenAI Whisper
``transcribe()`` / ``DecodingTask`` it is a port of) and anchors on the reference's call site: only ``result['text']`` is read.
  * 30 s windows slide over the audio; each is decoded WITH timestamp tokens and the seek advances to the last timestamp;
  * the previous window's text conditions the next one (``<|startofprev|>`` prompt) unless a fallback raised the temperature;
  * temperature fallback (0.0, 0.2, ... 1.0) on compression ratio > 2.4 or average log-probability < -1.0;
  * a window is skipped as silence when P(<|nospeech|>) > 0.6 and the average log-probability is below -1.0.
Everything numeric runs on the GPU: log-mel, encoder, and one ``wipa_decode_next`` per token through the stepwise C ABI,
with the logit filters (suppress lists, blank suppression, timestamp rules) and the sampling applied to the device logits.

Difference from the reference for audio LONGER than 30 s (none of its evaluation clips are): every window is featurized on
its own, so the reflect padding at interior window edges and the ``max - 8 dB`` clamp see the window instead of the whole file.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .audio import N_SAMPLES, SAMPLE_RATE, load_audio, log_mel_features
from . import decoding

HOP = 160
N_FRAMES = 3000
FRAMES_PER_SECOND = SAMPLE_RATE // HOP               # 100 mel frames per second
INPUT_STRIDE = 2                                     # mel frames per encoder position
TIME_PRECISION = INPUT_STRIDE * HOP / SAMPLE_RATE    # 0.02 s per timestamp token
N_TEXT_CTX = 448


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b)) if b else 0.0


@dataclass
class WindowResult:
    tokens: List[int]
    avg_logprob: float
    no_speech_prob: float
    temperature: float
    compression_ratio: float
    text: str = ""


def apply_timestamp_rules(logits: torch.Tensor, sampled: Sequence[int], arch, first: bool,
                          max_initial_timestamp_index: Optional[int] = 50) -> torch.Tensor:
    """Whisper's ApplyTimestampRules for one row (logits f32 [V] on the device, `sampled` = tokens after the prompt):
    timestamps come in pairs, never decrease, the first sampled token is a timestamp <= 1.0 s, and when the timestamps'
    total probability beats every text token a timestamp is forced."""
    tb, eot = arch.timestamp_begin, arch.eot
    logits[arch.no_timestamps] = -float("inf")
    seq = list(sampled)
    last_ts = len(seq) >= 1 and seq[-1] >= tb
    penult_ts = len(seq) < 2 or seq[-2] >= tb
    if last_ts:
        if penult_ts:
            logits[tb:] = -float("inf")              # a pair just closed: text (or EOT) must follow
        else:
            logits[:eot] = -float("inf")             # a segment just ended: a timestamp (or EOT) must follow
    ts = [t for t in seq if t >= tb]
    if ts:
        last = ts[-1] if (last_ts and not penult_ts) else ts[-1] + 1
        logits[tb:last] = -float("inf")              # timestamps do not decrease
    if first:
        logits[:tb] = -float("inf")
        if max_initial_timestamp_index is not None:
            logits[tb + max_initial_timestamp_index + 1:] = -float("inf")
    logprobs = torch.log_softmax(logits.float(), dim=-1)
    if torch.logsumexp(logprobs[tb:], dim=-1) > logprobs[:tb].max():
        logits[:tb] = -float("inf")
    return logits


def decode_window(model, prompt_tokens: Sequence[int], language: str, task: str, temperature: float,
                  without_timestamps: bool = False, suppress_tokens="-1", sample_len: Optional[int] = None,
                  generator: Optional[torch.Generator] = None) -> WindowResult:
    """One DecodingTask over the window whose encoder output the model holds (batch of 1)."""
    arch = model.arch
    sot_seq = [arch.sot, arch.language_token(language), arch.transcribe if task == "transcribe" else arch.translate]
    if without_timestamps:
        sot_seq.append(arch.no_timestamps)
    initial: List[int] = []
    if prompt_tokens:
        initial = [arch.sot_prev] + list(prompt_tokens)[-(N_TEXT_CTX // 2 - 1):]
    sot_index = len(initial)
    initial = initial + sot_seq
    sample_begin = len(initial)
    sample_len = sample_len or N_TEXT_CTX // 2
    suppress = torch.tensor(arch.resolve_suppress_tokens(suppress_tokens), dtype=torch.int64, device=model.device)
    blank = torch.tensor([220, arch.eot], dtype=torch.int64, device=model.device)

    # no-speech probability: softmax at the <|startoftranscript|> position
    if sot_index == 0:
        no_speech_prob = float(torch.softmax(model.decode_begin([[arch.sot]]).float(), dim=-1)[0, arch.no_speech])
    else:
        no_speech_prob = float(torch.softmax(model.decode_begin([initial[:sot_index + 1]]).float(), dim=-1)[0, arch.no_speech])
    logits = model.decode_begin([initial])[0]
    sampled: List[int] = []
    sum_logprob = 0.0
    for i in range(sample_len):
        if i == 0:
            logits[blank] = -float("inf")
        if suppress.numel():
            logits[suppress] = -float("inf")
        if not without_timestamps:
            logits = apply_timestamp_rules(logits, sampled, arch, first=(i == 0))
        logprobs = torch.log_softmax(logits.float(), dim=-1)
        if temperature == 0.0:
            nxt = int(logits.argmax())
        else:
            nxt = int(torch.multinomial(torch.softmax(logits.float() / temperature, dim=-1), 1, generator=generator))
        sum_logprob += float(logprobs[nxt])
        if nxt == arch.eot or len(initial) + len(sampled) + 1 >= N_TEXT_CTX:
            break
        sampled.append(nxt)
        if i + 1 < sample_len:
            logits = model.decode_next(torch.tensor([nxt], device=model.device))[0]
    avg = sum_logprob / (len(sampled) + 1)
    text = decoding._text([t for t in sampled if t < arch.eot]).strip()
    return WindowResult(sampled, avg, no_speech_prob, temperature, compression_ratio(text), text)


def decode_with_fallback(model, prompt_tokens, language, task, temperatures: Sequence[float], compression_ratio_threshold,
                         logprob_threshold, no_speech_threshold, generator=None, **kw) -> WindowResult:
    result = None
    for t in temperatures:
        result = decode_window(model, prompt_tokens, language, task, float(t), generator=generator, **kw)
        needs_fallback = False
        if compression_ratio_threshold is not None and result.compression_ratio > compression_ratio_threshold:
            needs_fallback = True                    # too repetitive
        if logprob_threshold is not None and result.avg_logprob < logprob_threshold:
            needs_fallback = True                    # average log probability is too low
        if no_speech_threshold is not None and result.no_speech_prob > no_speech_threshold:
            needs_fallback = False                   # silence
        if not needs_fallback:
            break
    return result


def transcribe(audio: Union[str, np.ndarray, torch.Tensor], model, *, language: Optional[str] = None, task: str = "transcribe",
               temperature: Union[float, Sequence[float]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
               compression_ratio_threshold: Optional[float] = 2.4, logprob_threshold: Optional[float] = -1.0,
               no_speech_threshold: Optional[float] = 0.6, condition_on_previous_text: bool = True,
               word_timestamps: bool = False, without_timestamps: bool = False, suppress_tokens="-1", seed: int = 0) -> Dict:
    """-> {"text", "segments", "language"} like mlx_whisper.transcribe.  `model` is a WhisperIPA (any max_batch >= 1)."""
    if word_timestamps:
        raise NotImplementedError("word_timestamps=True is outside the reference's path (it passes False)")
    arch = model.arch
    if isinstance(audio, str):
        audio = load_audio(audio)
    audio = np.asarray(audio.cpu() if isinstance(audio, torch.Tensor) else audio, dtype=np.float32).reshape(-1)
    content_frames = len(audio) // HOP
    temperatures = (temperature,) if isinstance(temperature, (int, float)) else tuple(temperature)
    gen = torch.Generator(device=model.device).manual_seed(seed)

    def window_features(seek: int):
        s0 = seek * HOP
        seg = audio[s0:s0 + N_SAMPLES]
        clip = np.zeros(N_SAMPLES, np.float32)
        clip[:len(seg)] = seg
        model.encoder(log_mel_features(clip[None], arch.n_mels), return_features=False)

    if language is None:
        window_features(0)
        language = decoding._detect_cached(model, 1)[0][0]

    seek = 0
    all_tokens: List[int] = []
    prompt_reset_since = 0
    segments: List[Dict] = []
    tb = arch.timestamp_begin
    while seek < content_frames:
        seek_before = seek
        time_offset = seek * HOP / SAMPLE_RATE
        segment_size = min(N_FRAMES, content_frames - seek)
        segment_duration = segment_size * HOP / SAMPLE_RATE
        window_features(seek)
        result = decode_with_fallback(model, all_tokens[prompt_reset_since:], language, task, temperatures,
                                      compression_ratio_threshold, logprob_threshold, no_speech_threshold, generator=gen,
                                      without_timestamps=without_timestamps, suppress_tokens=suppress_tokens)
        tokens = result.tokens
        if no_speech_threshold is not None:
            should_skip = result.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                should_skip = False                  # confident enough despite the no-speech probability
            if should_skip:
                seek += segment_size
                continue
        current: List[Dict] = []

        def new_segment(start, end, toks):
            text_tokens = [t for t in toks if t < arch.eot]
            current.append({"seek": seek, "start": start, "end": end, "tokens": list(toks), "text": decoding._text(text_tokens),
                            "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                            "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob})

        is_ts = [t >= tb for t in tokens]
        single_timestamp_ending = len(tokens) >= 2 and (not is_ts[-2]) and is_ts[-1]
        if len(tokens) == 1 and is_ts[-1]:
            single_timestamp_ending = True
        consecutive = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
        if consecutive:
            slices = list(consecutive)
            if single_timestamp_ending:
                slices.append(len(tokens))
            last_slice = 0
            for cur in slices:
                sl = tokens[last_slice:cur]
                start_pos, end_pos = sl[0] - tb, sl[-1] - tb
                new_segment(time_offset + start_pos * TIME_PRECISION, time_offset + end_pos * TIME_PRECISION, sl)
                last_slice = cur
            if single_timestamp_ending:
                seek += segment_size                 # no speech after the last timestamp
            else:
                seek += (tokens[last_slice - 1] - tb) * INPUT_STRIDE     # otherwise resume at the last timestamp
        else:
            duration = segment_duration
            ts = [t for t in tokens if t >= tb]
            if ts and ts[-1] != tb:
                duration = (ts[-1] - tb) * TIME_PRECISION
            new_segment(time_offset, time_offset + duration, tokens)
            seek += segment_size
        if seek <= seek_before:
            seek = seek_before + segment_size        # a closing timestamp of 0.00 must not stall the loop
        if not condition_on_previous_text or result.temperature > 0.5:
            prompt_reset_since = len(all_tokens)     # do not feed a prompt that is likely wrong
        for seg in current:
            if seg["start"] == seg["end"] or seg["text"].strip() == "":
                seg["text"], seg["tokens"] = "", []
        segments.extend(current)
        all_tokens.extend(t for seg in current for t in seg["tokens"])
    text = decoding._text([t for t in all_tokens if t < arch.eot])
    return {"text": text, "segments": segments, "language": language}
