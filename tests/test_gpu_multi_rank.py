"""The multi-GPU path that ships - Transcriber.evaluate_ids under NCCL, one process per GPU - against a single-process run
over the same utterances.  Needs >= 2 GPUs (skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi_rank.py -m gpu`)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, ws, port, n, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    import whisper_ipa_b200 as w
    from whisper_ipa_b200 import pipeline
    from oracle import hf_reference as hf
    from oracle import whisper_oracle as wo
    sd = hf.state_dict_f32(hf.build_hf_model("tiny", seed=0, init_gain=3.0))
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=3)
    m.load_state_dict(sd)
    audio = wo.synthetic_audio(n) * np.linspace(0.2, 2.0, n, dtype=np.float32)[:, None]        # every clip decodes differently
    refs = wo.synthetic_references(n)
    out = pipeline.Transcriber(m, max_new=12).evaluate_ids(audio, refs, micro_batch=3)          # full inputs, strided shard
    # the same sweep with only the local shard handed in (what a sharded loader provides)
    mine = out["local_indices"]
    out2 = pipeline.Transcriber(m, max_new=12).evaluate_ids(torch.from_numpy(audio[mine]).pin_memory(), refs, micro_batch=3,
                                                            local_shard=True)
    q.put((rank, out["counts"].tolist(), out["per_scores"], float(out["per"]), float(out["per_std"]), mine,
           out["local_hypotheses"], out2["per_scores"] == out["per_scores"]))
    m.close()
    dist.destroy_process_group()


def test_two_rank_nccl_evaluate_ids_matches_single_process(built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import whisper_ipa_b200 as w
    from whisper_ipa_b200 import pipeline
    from oracle import hf_reference as hf
    from oracle import per_oracle as po
    from oracle import whisper_oracle as wo
    n = 7                                                     # ragged: 4 utterances on rank 0, 3 on rank 1, padded gather
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=600) for _ in procs)
    [p.join(120) for p in procs]
    # single process over the same 7 utterances
    sd = hf.state_dict_f32(hf.build_hf_model("tiny", seed=0, init_gain=3.0))
    m = w.WhisperIPA("tiny", dtype="float32", max_batch=3)
    m.load_state_dict(sd)
    audio = wo.synthetic_audio(n) * np.linspace(0.2, 2.0, n, dtype=np.float32)[:, None]
    refs = wo.synthetic_references(n)
    want = pipeline.Transcriber(m, max_new=12).evaluate_ids(audio, refs, micro_batch=3)
    m.close()
    hyps = {}
    for rank, counts, per_scores, per, std, mine, local_hyps, same2 in got:
        assert counts == want["counts"].tolist() and per_scores == want["per_scores"]
        assert per == float(want["per"]) and std == float(want["per_std"]) and same2
        assert mine == list(range(rank, n, 2))
        hyps.update(dict(zip(mine, local_hyps)))
    assert [hyps[i] for i in range(n)] == want["local_hypotheses"]
    d = po.levenshtein_batch(refs, [np.asarray(hyps[i], np.int32) for i in range(n)])
    assert [int(x) for x in d] == [c[0] for c in want["counts"].tolist()]
