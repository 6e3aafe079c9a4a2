"""The batched transcription + evaluation hot path: audio -> log-mel -> encoder -> greedy decode -> PER, micro-batched on
one GPU and sharded across GPUs.  It is what ``evaluate_model``'s per-sample loop (ref:scripts/evaluate_model.py:179-209)
followed by ``evaluate_batch`` (ref:scripts/evaluate_ipa.py:346-378) computes, with B utterances per launch instead of 1."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import metrics, parallel
from .audio import log_mel_features
from .model import WhisperIPA


def hyps_to_csr(ids: torch.Tensor, lens: torch.Tensor):
    """Device ids [B, max_new] + lengths [B] -> (flat int32, offsets int32 [B+1]) without leaving the GPU."""
    B, L = ids.shape
    keep = torch.arange(L, device=ids.device)[None, :] < lens[:, None]
    flat = ids[keep].to(torch.int32).contiguous()
    off = torch.zeros(B + 1, dtype=torch.int32, device=ids.device)
    off[1:] = torch.cumsum(lens, 0)
    if flat.numel() == 0:
        flat = torch.zeros(1, dtype=torch.int32, device=ids.device)
    return flat, off


class Transcriber:
    def __init__(self, model: WhisperIPA, max_new: Optional[int] = None, prompt: Optional[Sequence[int]] = None,
                 num_beams: int = 1, length_penalty: float = 1.0):
        self.model = model
        self.num_beams, self.length_penalty = int(num_beams), float(length_penalty)
        self.prompt = list(prompt) if prompt is not None else model.arch.prompt("en", "transcribe", True)
        # the reference samples sample_len = 224 tokens after the prompt (mlx_whisper: `for i in range(sample_len)`)
        self.max_new = int(max_new) if max_new is not None else min(224, 448 - len(self.prompt))

    def transcribe_device(self, audio_dev: torch.Tensor):
        """audio f32 [B, 480000] already in HBM -> device (ids [B, max_new], lens [B])."""
        mel = log_mel_features(audio_dev, self.model.arch.n_mels)
        self.model.encoder(mel, return_features=False)
        return self.model.decode_tokens(self.prompt, self.max_new, num_beams=self.num_beams, length_penalty=self.length_penalty)

    def score_device(self, ids: torch.Tensor, lens: torch.Tensor, ref_flat: torch.Tensor, ref_off: torch.Tensor,
                     max_ref_len: int) -> torch.Tensor:
        hyp_flat, hyp_off = hyps_to_csr(ids, lens)
        return metrics.edit_distance_counts_device(ref_flat, ref_off, hyp_flat, hyp_off, max_ref_len)

    def evaluate_local(self, audio, references: Sequence[Sequence[int]], indices: Optional[Sequence[int]] = None,
                       micro_batch: Optional[int] = None, audio_rows: Optional[Sequence[int]] = None):
        """Transcribe + score the utterances `indices` of `audio` (host numpy / torch f32 [N, 480000], ideally pinned; or a
        device tensor) against `references[i]`.  Micro-batches are double-buffered: while batch k is being transcribed,
        batch k+1 is gathered into a pinned staging buffer and copied host->device on a side stream.
        ``audio_rows[j]`` names the row of `audio` that holds utterance `indices[j]` (default: the index itself).
        Returns (counts int32 [n, 2] on the device, hypotheses list, hypothesis lengths list), in `indices` order."""
        idx_all = list(range(len(references))) if indices is None else list(indices)
        rows_all = idx_all if audio_rows is None else list(audio_rows)
        if len(rows_all) != len(idx_all):
            raise ValueError("audio_rows must name one row per utterance")
        mb = micro_batch or self.model.max_batch
        dev = self.model.device
        n = len(idx_all)
        counts = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        on_device = isinstance(audio, torch.Tensor) and audio.is_cuda
        host = None if on_device else (audio if isinstance(audio, torch.Tensor) else torch.from_numpy(np.asarray(audio, dtype=np.float32)))
        copy_stream = torch.cuda.Stream(device=dev) if not on_device else None
        staging = [None, None]
        chunks = [idx_all[s:s + mb] for s in range(0, n, mb)]
        row_chunks = [rows_all[s:s + mb] for s in range(0, n, mb)]

        def stage(k):
            """Start the host->device copy of micro-batch k; returns (device tensor, ready event)."""
            idx = row_chunks[k]
            contiguous = all(idx[j] + 1 == idx[j + 1] for j in range(len(idx) - 1))
            if on_device:
                return (audio[idx[0]:idx[0] + len(idx)] if contiguous else audio[torch.as_tensor(idx, device=audio.device)]), None
            if contiguous and host.is_pinned():
                src = host[idx[0]:idx[0] + len(idx)]                      # a view: still pinned, no host copy
            else:
                if staging[k & 1] is None or staging[k & 1].shape[0] < len(idx):
                    staging[k & 1] = torch.empty((mb, host.shape[1]), dtype=torch.float32).pin_memory()
                src = staging[k & 1][:len(idx)]
                torch.index_select(host, 0, torch.as_tensor(idx), out=src)
            with torch.cuda.stream(copy_stream):
                d = src.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return d, ev

        results = []
        nxt = stage(0) if chunks else None
        for k, idx in enumerate(chunks):
            chunk, ev = nxt
            nxt = stage(k + 1) if k + 1 < len(chunks) else None           # overlaps with the transcription below
            if ev is not None:
                torch.cuda.current_stream(dev).wait_event(ev)
            ids, lens = self.transcribe_device(chunk)
            if ev is not None:
                chunk.record_stream(torch.cuda.current_stream(dev))
            rf, ro = metrics._pack([references[i] for i in idx])
            rf_d = torch.from_numpy(rf).to(dev, non_blocking=True)
            ro_d = torch.from_numpy(ro).to(dev, non_blocking=True)
            s0 = k * mb
            counts[s0:s0 + len(idx)] = self.score_device(ids, lens, rf_d, ro_d, int(np.max(np.diff(ro))) if len(idx) else 0)
            results.append((ids, lens))
        hyps: List[List[int]] = []
        hyp_lens: List[int] = []
        for ids, lens in results:                                         # one read-back pass after everything is queued
            ids_h, lens_h = ids.cpu().numpy(), lens.cpu().numpy()
            for j in range(ids_h.shape[0]):
                hyps.append(ids_h[j, :lens_h[j]].tolist())
                hyp_lens.append(int(lens_h[j]))
        return counts, hyps, hyp_lens

    def evaluate_ids(self, audio, references: Sequence[Sequence[int]], micro_batch: Optional[int] = None,
                     local_shard: bool = False, local_rows: Optional[Sequence[int]] = None) -> Dict:
        """audio: host (numpy / pinned torch) or device f32 [N, 480000]; references: N id sequences (always the full list).
        Every rank works on its strided shard `i % world == rank` and the per-utterance (distance, reference length) pairs are
        exchanged ONCE per sweep; returns the evaluate_batch-style PER dict (identical on every rank) plus the hypotheses of
        the local shard.  ``local_shard``: `audio` holds only this rank's clips, in shard order (row j = utterance
        rank + j * world) - what a sharded data loader provides - instead of all N; ``local_rows[j]`` then overrides which row
        of `audio` local utterance j uses (a sweep that re-uses a resident set of clips)."""
        n_total = len(references)
        rank, ws = parallel.world()
        mine = parallel.shard_indices(n_total, rank, ws)
        dev = self.model.device
        if local_shard:
            rows = list(range(len(mine))) if local_rows is None else list(local_rows)
            if len(rows) != len(mine) or (len(rows) and max(rows) >= len(audio)):
                raise ValueError(f"local_shard: {len(rows)} audio rows (max {max(rows) if rows else -1}) for a shard of {len(mine)} "
                                 f"utterances and {len(audio)} clips")
            counts_local, hyps_local, lens_local = self.evaluate_local(audio, references, mine, micro_batch, audio_rows=rows)
        else:
            counts_local, hyps_local, lens_local = self.evaluate_local(audio, references, mine, micro_batch)
        hyp_lens = np.zeros(n_total, dtype=np.int64)
        hyp_lens[mine] = lens_local
        table = parallel.gather_counts(counts_local, n_total)
        if ws > 1:
            import torch.distributed as dist
            hl = torch.from_numpy(hyp_lens).to(dev)
            dist.all_reduce(hl)                      # each utterance is owned by exactly one rank
            hyp_lens = hl.cpu().numpy()
        per = [metrics.per_from_counts(int(table[i, 0]), int(table[i, 1]), int(hyp_lens[i])) for i in range(n_total)]
        out = metrics.summarize(per)
        out.update({"counts": table, "local_indices": mine, "local_hypotheses": hyps_local,
                    "sum_edits": int(table[:, 0].astype(np.int64).sum()), "sum_ref_len": int(table[:, 1].astype(np.int64).sum())})
        return out
