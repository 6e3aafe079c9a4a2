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
    def __init__(self, model: WhisperIPA, max_new: Optional[int] = None, prompt: Optional[Sequence[int]] = None):
        self.model = model
        self.prompt = list(prompt) if prompt is not None else model.arch.prompt("en", "transcribe", True)
        self.max_new = int(max_new) if max_new is not None else 224 - len(self.prompt)

    def transcribe_device(self, audio_dev: torch.Tensor):
        """audio f32 [B, 480000] already in HBM -> device (ids [B, max_new], lens [B])."""
        mel = log_mel_features(audio_dev, self.model.arch.n_mels)
        self.model.encoder(mel, return_features=False)
        return self.model.decode_tokens(self.prompt, self.max_new)

    def score_device(self, ids: torch.Tensor, lens: torch.Tensor, ref_flat: torch.Tensor, ref_off: torch.Tensor,
                     max_ref_len: int) -> torch.Tensor:
        hyp_flat, hyp_off = hyps_to_csr(ids, lens)
        return metrics.edit_distance_counts_device(ref_flat, ref_off, hyp_flat, hyp_off, max_ref_len)

    def evaluate_ids(self, audio, references: Sequence[Sequence[int]], micro_batch: Optional[int] = None) -> Dict:
        """audio: host (numpy / pinned torch) or device f32 [N, 480000]; references: N id sequences.
        Every rank passes the FULL inputs and works on its strided shard; returns the evaluate_batch-style PER dict
        (identical on every rank) plus the hypotheses of the local shard."""
        n_total = len(references)
        rank, ws = parallel.world()
        mine = parallel.shard_indices(n_total, rank, ws)
        mb = micro_batch or self.model.max_batch
        dev = self.model.device
        counts_local = torch.zeros((len(mine), 2), dtype=torch.int32, device=dev)
        hyps_local: List[List[int]] = []
        hyp_lens = np.zeros(n_total, dtype=np.int64)
        for s in range(0, len(mine), mb):
            idx = mine[s:s + mb]
            chunk = audio[idx] if not isinstance(audio, torch.Tensor) else audio[torch.as_tensor(idx)]
            if isinstance(chunk, np.ndarray):
                chunk = torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32)).pin_memory()
            chunk = chunk.to(dev, non_blocking=True)
            ids, lens = self.transcribe_device(chunk)
            refs = [references[i] for i in idx]
            rf, ro = metrics._pack(refs)
            rf_d = torch.from_numpy(rf).to(dev)
            ro_d = torch.from_numpy(ro).to(dev)
            counts_local[s:s + len(idx)] = self.score_device(ids, lens, rf_d, ro_d, int(np.max(np.diff(ro))) if len(idx) else 0)
            ids_h, lens_h = ids.cpu().numpy(), lens.cpu().numpy()
            for j, i in enumerate(idx):
                hyps_local.append(ids_h[j, :lens_h[j]].tolist())
                hyp_lens[i] = lens_h[j]
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
