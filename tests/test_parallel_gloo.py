"""N>1 host path on CPU: two gloo ranks shard an utterance list, score their shards with the oracle as a stand-in for
the kernel, all-gather the (distance, ref_len) pairs and must reproduce the single-process aggregate bit for bit."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_ipa_b200 import metrics, parallel


def _pairs(n):
    rng = np.random.default_rng(11)
    refs = [rng.integers(0, 30, size=rng.integers(1, 60)).astype(np.int32) for _ in range(n)]
    hyps = [rng.integers(0, 30, size=rng.integers(0, 60)).astype(np.int32) for _ in range(n)]
    return refs, hyps


def _worker(rank, ws, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from oracle import per_oracle as po
    refs, hyps = _pairs(n)
    mine = parallel.shard_indices(n, rank, ws)
    local = torch.tensor([[po.levenshtein(refs[i], hyps[i]), len(refs[i])] for i in mine], dtype=torch.int32).reshape(-1, 2)
    table = parallel.gather_counts(local, n)
    per = [metrics.per_from_counts(int(table[i, 0]), int(table[i, 1]), len(hyps[i])) for i in range(n)]
    q.put((rank, table.tolist(), metrics.summarize(per)["per"], metrics.summarize(per)["per_std"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 16])
def test_two_rank_gather_matches_single_process(n):
    from oracle import per_oracle as po
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    refs, hyps = _pairs(n)
    want_counts = [[po.levenshtein(r, h), len(r)] for r, h in zip(refs, hyps)]
    want = metrics.summarize([metrics.per_from_counts(d, l, len(h)) for (d, l), h in zip(want_counts, hyps)])
    for rank, table, per, std in got:
        assert table == want_counts
        assert per == want["per"] and std == want["per_std"]
