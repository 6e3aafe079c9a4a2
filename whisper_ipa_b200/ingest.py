"""Audio ingest for the batched path: a worker pool reads / decodes files while the GPU transcribes the previous
micro-batch, raw PCM16 goes host -> device through pinned staging buffers on a side stream, and ONE kernel per source
format (csrc/resample.cu) does what the reference does per file on the host: s16le -> float / 32768 -> mono -> 16 kHz
(polyphase FIR, scipy.signal.resample_poly's) -> pad_or_trim to 30 s
(ref:scripts/evaluate_model.py:187-188, ref:scripts/transcribe_single.py:43-44, ref:scripts/ipa_data_loader.py:48,80).

Formats: PCM16 WAV at any rate / channel count takes the GPU path.  Anything else is decoded the way mlx_whisper.load_audio
does it - ``ffmpeg -i file -f s16le -ac 1 -ar 16000 -`` in a worker - when an ffmpeg binary exists (its output then takes
the same GPU path with rate 16000, 1 channel); 8- / 32-bit PCM WAV without ffmpeg goes through audio.load_audio on the
host.  A file that cannot be read yields a row of silence and its exception in ``errors`` (the reference's per-sample
``except`` turns it into an empty hypothesis, ref:scripts/evaluate_model.py:202-204).
"""
from __future__ import annotations

import math
import os
import shutil
import subprocess
import wave
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .audio import N_SAMPLES, SAMPLE_RATE, load_audio


@dataclass(frozen=True)
class ResamplePlan:
    up: int
    down: int
    taps: np.ndarray          # float32 [n_taps] = up * firwin(...)
    c0: int                   # out[n] = sum_i taps[(n * down + c0) - i * up] * x[i]

    def frames_needed(self, n_out: int) -> int:
        """Source frames that can influence the first n_out output samples."""
        return ((n_out - 1) * self.down + self.c0) // self.up + 2


@lru_cache(maxsize=None)
def resample_plan(rate: int, sr: int = SAMPLE_RATE) -> ResamplePlan:
    """Filter and index bookkeeping of ``scipy.signal.resample_poly(x, sr / g, rate / g)`` (default Kaiser-5.0 window):
    h = up * firwin(2 * half_len + 1, 1 / max(up, down)), half_len = 10 * max(up, down); scipy prepends n_pre_pad zeros to h
    and drops the first n_pre_remove outputs of upfirdn, which folds into c0 = n_pre_remove * down - n_pre_pad."""
    g = math.gcd(int(rate), int(sr))
    up, down = sr // g, rate // g
    if up == 1 and down == 1:
        return ResamplePlan(1, 1, np.ones(1, np.float32), 0)
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    m = np.arange(n) - 0.5 * (n - 1)
    h = f_c * np.sinc(f_c * m) * np.kaiser(n, 5.0)       # firwin(n, f_c, window=("kaiser", 5.0)): windowed sinc, unit DC gain
    h = h / h.sum() * up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return ResamplePlan(up, down, h.astype(np.float32), n_pre_remove * down - n_pre_pad)


def _ffmpeg() -> Optional[str]:
    return shutil.which("ffmpeg")


def read_pcm(path: str, max_seconds: float = 30.0, pin=None) -> Dict:
    """One file -> {"pcm": int16 [frames * ch], "frames", "ch", "rate"} (the GPU path) or {"f32": float32 16 kHz mono}.
    ``pin(n_elems)`` (optional) hands out a pinned int16 torch buffer: the samples are then copied into it here, in the
    worker thread, and "pcm" is that buffer (host->device copies start from it without another pass over the data)."""
    def land(pcm):
        if pin is None:
            return pcm
        buf = pin(pcm.size)
        np.copyto(buf.numpy()[:pcm.size], pcm)
        return buf
    try:
        with wave.open(path, "rb") as w:
            ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            comp = w.getcomptype()
            if width == 2 and comp == "NONE" and 1 <= ch <= 8:
                need = resample_plan(rate).frames_needed(int(max_seconds * SAMPLE_RATE))
                n = min(n, need)
                raw = w.readframes(n)
                pcm = np.frombuffer(raw, dtype="<i2")
                return {"pcm": land(pcm), "frames": len(pcm) // ch, "ch": ch, "rate": rate}
    except (wave.Error, EOFError):
        pass                                               # not a RIFF/WAVE file: try ffmpeg below
    exe = _ffmpeg()
    if exe is not None:
        # exactly mlx_whisper.audio.load_audio's command: decode, down-mix and resample in ffmpeg, s16le on stdout
        cmd = [exe, "-nostdin", "-threads", "0", "-i", path, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le", "-ar",
               str(SAMPLE_RATE), "-"]
        out = subprocess.run(cmd, capture_output=True, check=True).stdout
        pcm = np.frombuffer(out, dtype="<i2")[: int(max_seconds * SAMPLE_RATE)]
        return {"pcm": land(pcm), "frames": len(pcm), "ch": 1, "rate": SAMPLE_RATE}
    return {"f32": load_audio(path)[: int(max_seconds * SAMPLE_RATE)]}     # 8- / 32-bit PCM WAV; raises on other containers


class _PinnedPool:
    """Reusable pinned int16 buffers in size classes of 256 Ki samples (pinning is slow, so buffers are recycled; handing one out
    and taking it back are thread-safe: the pool's workers call get(), the batch loader calls put() once the host->device
    copy that reads the buffer has been enqueued AND has finished)."""

    def __init__(self):
        import threading
        self._lock = threading.Lock()
        self._free: Dict[int, List[torch.Tensor]] = {}

    def get(self, n_elems: int) -> torch.Tensor:
        size = ((max(n_elems, 1) + 262143) // 262144) * 262144
        with self._lock:
            lst = self._free.get(size)
            if lst:
                return lst.pop()
        return torch.empty(size, dtype=torch.int16).pin_memory()

    def put(self, buf: torch.Tensor) -> None:
        with self._lock:
            self._free.setdefault(buf.numel(), []).append(buf)


class AudioIngest:
    """files -> device f32 [B, 480000], overlapped with whatever the GPU is doing.

    Worker threads read each file and land its samples in a pinned buffer; the batch loader (the caller's thread for
    load_batch, a producer thread for iter_batches) enqueues one host->device copy per file and one resampler launch per
    (rate, channels) group on the ingest stream.  iter_batches keeps one batch in flight: while the caller transcribes batch
    k, batch k + 1 is read, copied and resampled."""

    def __init__(self, device=None, workers: Optional[int] = None, chunk: int = 64):
        if not torch.cuda.is_available():
            raise RuntimeError("whisper_ipa_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.workers = workers or min(32, os.cpu_count() or 4)
        self.chunk = int(chunk)
        self.pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="wipa-ingest")
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = _PinnedPool()
        self._taps: Dict[int, torch.Tensor] = {}
        self._pending: List[Tuple[torch.cuda.Event, List[torch.Tensor]]] = []    # pinned buffers still being read by a copy

    def close(self) -> None:
        self.pool.shutdown(wait=False, cancel_futures=True)

    # ---- host side -----------------------------------------------------------------------------
    def submit(self, paths: Sequence[str]):
        """Start reading `paths` in the pool; returns the futures (one per file) for load_batch(futures=...)."""
        return [self.pool.submit(read_pcm, p, 30.0, self._pinned.get) for p in paths]

    def _recycle(self, block: bool = False) -> None:
        keep = []
        for ev, bufs in self._pending:
            if block:
                ev.synchronize()
            if ev.query():
                for b in bufs:
                    self._pinned.put(b)
            else:
                keep.append((ev, bufs))
        self._pending = keep

    def _taps_dev(self, rate: int) -> torch.Tensor:
        if rate not in self._taps:
            self._taps[rate] = torch.from_numpy(resample_plan(rate).taps).to(self.device)
        return self._taps[rate]

    # ---- device side ---------------------------------------------------------------------------
    def load_batch(self, paths: Optional[Sequence[str]] = None, futures=None, wait: bool = True
                   ) -> Tuple[torch.Tensor, List[Optional[Exception]]]:
        """-> (audio f32 [B, 480000] on the device, errors[B]).  Work is enqueued on the ingest stream; with ``wait`` the CURRENT
        stream is made to wait for it before this returns, so the result can be consumed right away."""
        futs = futures if futures is not None else self.submit(paths)
        B = len(futs)
        errors: List[Optional[Exception]] = [None] * B
        lib = _lib.lib()
        self._recycle()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            out = torch.zeros((B, N_SAMPLES), dtype=torch.float32, device=self.device)
            for k0 in range(0, B, self.chunk):
                items = []
                for b in range(k0, min(B, k0 + self.chunk)):
                    try:
                        items.append((b, futs[b].result()))
                    except Exception as e:                                     # unreadable file: silence + the exception
                        errors[b] = e
                groups: Dict[Tuple[int, int], List[Tuple[int, Dict]]] = {}
                for b, it in items:
                    if "f32" in it:
                        x = torch.from_numpy(np.ascontiguousarray(it["f32"][:N_SAMPLES], dtype=np.float32))
                        out[b, : x.numel()].copy_(x.pin_memory(), non_blocking=True)
                    else:
                        groups.setdefault((it["rate"], it["ch"]), []).append((b, it))
                if not groups:
                    continue
                total = sum(it["frames"] * it["ch"] for g in groups.values() for _, it in g)
                dev_pcm = torch.empty(max(total, 1), dtype=torch.int16, device=self.device)
                pos, launches, used = 0, [], []
                for (rate, ch), g in groups.items():
                    offs, frames, rows = [], [], []
                    for b, it in g:
                        n = it["frames"] * ch
                        src = it["pcm"] if isinstance(it["pcm"], torch.Tensor) else torch.from_numpy(np.ascontiguousarray(it["pcm"])).pin_memory()
                        if n:
                            dev_pcm[pos:pos + n].copy_(src[:n], non_blocking=True)       # pinned -> device, one copy per file
                        used.append(src)
                        offs.append(pos)
                        frames.append(it["frames"])
                        rows.append(b)
                        pos += n
                    launches.append((rate, ch, offs, frames, rows))
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self._pending.append((ev, [u for u in used if u.is_pinned()]))
                for rate, ch, offs, frames, rows in launches:
                    plan = resample_plan(rate)
                    meta = torch.tensor(offs, dtype=torch.int64).pin_memory().to(self.device, non_blocking=True)
                    nfr = torch.tensor(frames, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
                    contiguous = rows == list(range(rows[0], rows[0] + len(rows)))
                    dst = out[rows[0]:rows[0] + len(rows)] if contiguous else torch.empty((len(rows), N_SAMPLES), dtype=torch.float32,
                                                                                         device=self.device)
                    taps = self._taps_dev(rate)
                    _lib.check(lib.wipa_resample_pcm16(dev_pcm.data_ptr(), meta.data_ptr(), nfr.data_ptr(), len(rows), ch, plan.up,
                                                       plan.down, taps.data_ptr(), taps.numel(), plan.c0, dst.data_ptr(), N_SAMPLES,
                                                       self.stream.cuda_stream), "wipa_resample_pcm16")
                    if not contiguous:
                        out.index_copy_(0, torch.tensor(rows, device=self.device), dst)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._last_done = done
        if wait:
            torch.cuda.current_stream(self.device).wait_event(done)
            out.record_stream(torch.cuda.current_stream(self.device))
        return out, errors

    def iter_batches(self, paths: Sequence[str], batch_size: int) -> Iterator[Tuple[List[int], torch.Tensor, List[Optional[Exception]]]]:
        """Yield (indices, audio [b, 480000] on the device, errors) per micro-batch.  A producer thread stays one batch ahead:
        while the caller works on batch k, the files of batch k + 1 are read by the pool, copied host->device and resampled on
        the ingest stream (and the files of batch k + 2 are already being read)."""
        import queue
        import threading
        chunks = [list(range(s, min(len(paths), s + batch_size))) for s in range(0, len(paths), batch_size)]
        if not chunks:
            return
        q: "queue.Queue" = queue.Queue(maxsize=1)

        def producer():
            try:
                futs = self.submit([paths[i] for i in chunks[0]])
                for k, idx in enumerate(chunks):
                    nxt = self.submit([paths[i] for i in chunks[k + 1]]) if k + 1 < len(chunks) else None
                    audio, errors = self.load_batch(futures=futs, wait=False)
                    q.put((idx, audio, errors, self._last_done, None))
                    futs = nxt
            except BaseException as e:                       # surfaces in the consumer
                q.put((None, None, None, None, e))
        t = threading.Thread(target=producer, name="wipa-ingest-producer", daemon=True)
        t.start()
        cur = torch.cuda.current_stream(self.device)
        for _ in chunks:
            idx, audio, errors, done, exc = q.get()
            if exc is not None:
                raise exc
            cur.wait_event(done)
            audio.record_stream(cur)
            yield idx, audio, errors
        t.join()
        self._recycle(block=True)
