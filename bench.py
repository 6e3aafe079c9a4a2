#!/usr/bin/env python
"""Headline benchmark: whisper-small greedy IPA decode + PER, audio-seconds transcribed per second (RTFx).

One "step" = one pass of the whole hot path over one micro-batch of synthetic clips per GPU:
log-mel -> encoder -> 220 greedy decode steps (4-token prompt, 224 positions) -> PER counts.  A run of K steps is ONE
evaluation sweep of K * B * N utterances sharded over the N ranks (utterance i -> rank i mod N) through the public API,
``pipeline.Transcriber.evaluate_ids``: per-rank micro-batches, ONE all-gather of the (edit distance, reference length)
pairs per sweep, the aggregate PER finished with the reference's numpy expressions.

`value`    : the sweep with this rank's audio already resident in HBM (CUDA events, max over ranks).
`e2e`      : the same sweep from pinned HOST buffers: host->device copies (double-buffered on a side stream), reference
             upload, PER gather and the read-back of ids / lengths / counts inside the timed region; mean of the timed passes.
`roofline` : the dominant kernel (decoder cross-attention: a persistent HBM streamer) timed alone with CUDA events on the
             decode shapes (bytes per launch >> L2), achieved GB/s of its ALGORITHMIC bytes vs MEASURED_PEAKS.json; `traffic`
             is read from the committed ncu capture profiles/<kernel>.json and dropped when the kernel source changed since.
             `roofline.decode_step` / `roofline.logits_argmax`: the whole decode step and its vocabulary node the same way.
`parity`   : the benchmarked build against HF transformers fp32 on the same GPU (TF32 off): teacher-forced logits and
             220-step greedy ids for the first clips of this very workload.
`gpu_baseline`: HF transformers ``generate`` in bf16 / SDPA on the same B200 (the library path a user would otherwise run).
`cpu_baseline` / ``--impl reference``: the parity oracle (HF transformers Whisper on the host cores, fp32) on a bounded
             sample of the same workload.  The reference's own runtime (mlx_whisper on Apple Metal) cannot run here.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 30.0
N_SAMPLES = 480000
CONFIG_OF_ARCH = {"tiny": "configs[0]", "base": "configs[1]", "small": "configs[2], the headline", "medium": "configs[3]",
                  "large-v3": "configs[4]"}


def synthetic_audio(n, first=0):
    out = np.empty((n, N_SAMPLES), dtype=np.float32)
    for i in range(n):
        out[i] = np.random.default_rng(1234 + first + i).standard_normal(N_SAMPLES).astype(np.float32) * 0.1
    return out


def synthetic_references(n):
    rng = np.random.default_rng(4321)
    return [rng.integers(0, 50257, size=int(rng.integers(20, 120))).astype(np.int32) for _ in range(n)]


def random_init_state_dict(arch_name, seed=0):
    """Random-init weights of the named architecture (HF default init; there is no network for checkpoints)."""
    import logging
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from whisper_ipa_b200 import ARCHS
    logging.getLogger("transformers").setLevel(logging.ERROR)
    cfg = WhisperConfig(**ARCHS[arch_name].hf_config_kwargs(), decoder_start_token_id=50258, eos_token_id=50257,
                        pad_token_id=50257, bos_token_id=50257, begin_suppress_tokens=[220, 50257])
    torch.manual_seed(seed)
    model = WhisperForConditionalGeneration(cfg).eval()
    return model, {k: v.detach().float() for k, v in model.state_dict().items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(kernel, source_rel, shape_key):
    """dram__bytes_read + dram__bytes_write per launch of `kernel` from profiles/<kernel>.json (written by
    scripts/ncu_summary.py --json from an `ncu --set full` capture).  Returns (bytes or None, note): a capture taken from a
    different version of the kernel source, or on another shape, is refused."""
    p = os.path.join(ROOT, "profiles", f"{kernel}.json")
    if not os.path.exists(p):
        return None, f"no capture committed (profiles/{kernel}.json)"
    d = json.load(open(p))
    sha = hashlib.sha256(open(os.path.join(ROOT, source_rel), "rb").read()).hexdigest()[:16]
    if d.get("source_sha16") != sha:
        return None, f"profiles/{kernel}.json was captured from another version of {source_rel}: stale, not quoted"
    if d.get("shape") != shape_key:
        return None, f"profiles/{kernel}.json was captured on shape {d.get('shape')!r}, this run is {shape_key!r}"
    return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), f"profiles/{kernel}.json ({d.get('captured', '?')})"


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (HF transformers fp32) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_oracle_pass(hf_model, fe, audio, refs, prompt, max_new):
    """Features + generate + PER over the given clips on the host. Returns (seconds, mean PER)."""
    from oracle import per_oracle as po
    t0 = time.perf_counter()
    feats = fe(list(audio), sampling_rate=16000, return_tensors="pt").input_features
    p = torch.tensor([prompt] * len(audio))
    with torch.no_grad():
        ids = hf_model.generate(feats, decoder_input_ids=p, max_new_tokens=max_new, do_sample=False)
    d = po.levenshtein_batch(refs, [np.asarray(r, np.int32) for r in ids.tolist()])
    per = np.mean([po.per_from_counts(int(x), len(r), ids.shape[1]) for x, r in zip(d, refs)])
    return time.perf_counter() - t0, float(per)


def run_reference(args):
    """--impl reference: the reference's CPU arm = the HF oracle on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import logging
    import warnings
    from transformers import WhisperFeatureExtractor
    from whisper_ipa_b200 import ARCHS
    warnings.filterwarnings("ignore")
    logging.getLogger("transformers").setLevel(logging.ERROR)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = ARCHS[args.arch]
    hf_model, _ = random_init_state_dict(args.arch)
    fe = WhisperFeatureExtractor(feature_size=arch.n_mels)          # built once, outside the timed region
    prompt = arch.prompt("en", "transcribe", True)
    n_clips = args.ref_clips
    refs = synthetic_references(n_clips)
    clips = [synthetic_audio(n_clips, first=s * n_clips) for s in range(max(args.steps, 1))]    # synthesis is not timed either
    for _ in range(args.warmup):
        cpu_oracle_pass(hf_model, fe, clips[0], refs, prompt, args.max_new)
    t = 0.0
    for s in range(args.steps):
        dt, _ = cpu_oracle_pass(hf_model, fe, clips[s], refs, prompt, args.max_new)
        t += dt
    value = n_clips * CLIP_SECONDS * args.steps / t
    sample = (f"{n_clips} clips x {args.max_new} greedy tokens + PER per step (--ref-clips {n_clips}: a bounded sample of the "
              f"{args.batch}-clip step), HF transformers fp32 on {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper-{args.arch} greedy IPA decode + PER (random-init weights, synthetic 30 s clips, BASELINE "
                               f"{CONFIG_OF_ARCH[args.arch]}); reference arm: {n_clips} clips per step on the host cores",
                   "clips_per_step": n_clips, "max_new_tokens": args.max_new},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
# parity of the benchmarked build and the GPU comparator (rank 0, outside every timed region)
# ---------------------------------------------------------------------------------------------------------------------
def parity_vs_hf_fp32(model, hf_model, audio_host, prompt, max_new, hyps_first, n_clips=16, logit_steps=32):
    """The benchmarked build vs HF fp32 on this GPU (TF32 off), on the first clips of the benchmark's own workload."""
    from oracle import hf_reference as hf
    import whisper_ipa_b200 as w
    n = min(n_clips, audio_host.shape[0])
    audio = audio_host[:n].numpy()
    ref = hf.hf_gpu_fp32_reference(hf_model, audio, prompt, max_new, logit_steps)
    mel = w.log_mel_features(audio, model.arch.n_mels)
    mel_err = (mel.cpu() - ref["mel"]).abs().max().item()
    enc = model.encoder(mel).cpu()
    enc_rel = ((enc - ref["enc"]).norm() / ref["enc"].norm()).item()
    got = model.teacher_forced_logits(ref["tokens"]).cpu()
    rel = ((got - ref["logits"]).norm() / ref["logits"].norm()).item()
    P = len(prompt)
    tf_agree = (got[:, P - 1:].argmax(-1) == ref["logits"][:, P - 1:].argmax(-1)).float().mean().item()
    want = ref["ids"]
    L = want.shape[1]
    mine = torch.tensor([list(h[:L]) + [50257] * (L - len(h[:L])) for h in hyps_first[:n]])
    agree = (mine == want).float().mean().item()
    return {"oracle": "HF transformers fp32 on this GPU, TF32 off", "clips": n, "mel_max_abs": mel_err, "encoder_rel_l2": enc_rel,
            "logits_rel_l2": rel, "logit_positions": int(got.shape[1]), "teacher_forced_argmax_agreement": tf_agree,
            "token_agreement": agree, "tokens_compared": int(mine.numel()),
            "rows_identical": int((mine == want).all(dim=1).sum())}


def hf_gpu_baseline(hf_model, audio_host, prompt, max_new, batch, beams=1):
    """HF transformers generate() in bf16 with SDPA attention on this GPU: features on the GPU-resident model, same prompt,
    same token budget.  Largest batch that fits, capped at the benchmark's."""
    import warnings
    from transformers import WhisperFeatureExtractor
    dev = torch.device("cuda", torch.cuda.current_device())
    fe = WhisperFeatureExtractor(feature_size=hf_model.config.num_mel_bins)
    n = min(batch, audio_host.shape[0])
    feats = fe(list(audio_host[:n].numpy()), sampling_rate=16000, return_tensors="pt").input_features
    import copy
    m = copy.deepcopy(hf_model).to(dev).to(torch.bfloat16).eval()
    try:
        m.config._attn_implementation = "sdpa"
    except Exception:
        pass
    p = torch.tensor([prompt] * n, device=dev)
    kw = dict(decoder_input_ids=p, max_new_tokens=max_new, do_sample=False)
    if beams > 1:
        kw.update(num_beams=beams, length_penalty=1.0, early_stopping=False)
    out = None
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = feats.to(dev, torch.bfloat16)
            m.generate(f[:min(n, 8)], **{**kw, "decoder_input_ids": p[:min(n, 8)], "max_new_tokens": 8})      # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f = feats.pin_memory().to(dev, torch.bfloat16, non_blocking=True)      # host features in, like the HF pipeline
            ids = m.generate(f, **kw)
            ids.cpu()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        out = {"value": n * CLIP_SECONDS / (ms / 1000.0), "unit": "audio-s/s", "kind": "HF transformers generate, bf16, sdpa, on this GPU",
               "clips": n, "ms": ms, "note": "log-mel on the host is NOT timed (HF's extractor is numpy); encoder + generate + read-back are"}
    except Exception as e:                                     # e.g. out of memory at this batch: report, do not fail the bench
        out = {"value": None, "unit": "audio-s/s", "kind": "HF transformers generate, bf16, sdpa", "error": repr(e)[:200]}
    finally:
        del m
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="small", choices=list(CONFIG_OF_ARCH))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("WIPA_BENCH_BATCH", "512")),
                    help="clips per GPU per step = the micro-batch of the eval sweep (measured on a B200: 11.8k audio-s/s at 256, 12.5k at 384, 13.2k at 512)")
    ap.add_argument("--beams", type=int, default=1, help="beam search width (BASELINE configs[3]: medium, 5 beams)")
    ap.add_argument("--dtype", default="float16", choices=["float16", "bfloat16", "float32"],
                    help="float16 (default: libwipa.so, logits within 1e-3 of the fp32 oracle), bfloat16 (libwipa_bf16.so) or float32")
    ap.add_argument("--max-new", type=int, default=220)
    ap.add_argument("--ref-clips", type=int, default=8, help="clips per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-baseline-clips", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    import whisper_ipa_b200 as w
    from whisper_ipa_b200 import _lib, metrics, parallel, pipeline

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup
    arch = w.ARCHS[args.arch]

    hf_model, sd = random_init_state_dict(args.arch)
    model = w.WhisperIPA(args.arch, dtype=args.dtype, max_batch=B, max_beams=args.beams)
    model.load_state_dict(sd)
    del sd
    tr = pipeline.Transcriber(model, max_new=args.max_new, num_beams=args.beams)
    P = len(tr.prompt)

    # The sweep: K * B * world utterances; utterance i belongs to rank i % world and is its (i // world)-th local clip.
    # Every rank re-uses B distinct synthetic clips K times (weights and audio stay resident; no clip is cached across steps:
    # the whole path runs for each), references are distinct per utterance.
    n_total = K * B * world
    audio_b = torch.from_numpy(synthetic_audio(B, first=rank * B)).pin_memory()          # this rank's B resident clips (pinned)
    audio_dev = audio_b.to(dev)
    rows = [j % B for j in range(K * B)]                 # local utterance j -> clip j % B
    refs_all = synthetic_references(n_total)
    n_warm = B * world
    refs_warm = refs_all[:n_warm]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):                                   # warm-up steps: one micro-batch per rank each, same code path
        tr.evaluate_ids(audio_dev, refs_warm, micro_batch=B, local_shard=True)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    if args.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    res_dev = tr.evaluate_ids(audio_dev, refs_all, micro_batch=B, local_shard=True, local_rows=rows)      # K steps per rank, ONE gather
    e1.record()
    sync_all()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    launches = torch.tensor([_lib.launch_count()], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches)
    ms_total = float(ms.item())
    value = n_total * CLIP_SECONDS / (ms_total / 1000.0)

    # ---- e2e: pinned host audio in, host results out, through the same public call --------------------------------------
    tr.evaluate_ids(audio_b, refs_warm, micro_batch=B, local_shard=True)          # warm-up: staging, side stream
    sync_all()
    e2e_ms = []
    for _ in range(2):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        res_e2e = tr.evaluate_ids(audio_b, refs_all, micro_batch=B, local_shard=True, local_rows=rows)
        t1.record()
        sync_all()
        e2e_ms.append(t0.elapsed_time(t1))
    ms2 = torch.tensor([float(np.mean(e2e_ms))], device=dev)             # mean of the timed passes
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = n_total * CLIP_SECONDS / (float(ms2.item()) / 1000.0)
    mean_ref = float(np.mean([len(r) for r in refs_all]))
    h2d = B * N_SAMPLES * 4 + int(B * mean_ref * 4) + (B + 1) * 4
    d2h = B * args.max_new * 4 + B * 4 + B * 8
    clocks = sampler.stop() if rank == 0 else None

    # ---- the sharded aggregate equals a single-process aggregate over the same utterances ---------------------------------
    # every rank's hypotheses go to rank 0 (outside the timed regions), which scores all of them with the CPU oracle
    mine = res_dev["local_indices"]
    hyp_pad = torch.full((len(mine), args.max_new), -1, dtype=torch.int32)
    for j, h in enumerate(res_dev["local_hypotheses"]):
        hyp_pad[j, :len(h)] = torch.tensor(h, dtype=torch.int32)
    if world > 1:
        gathered = [torch.empty_like(hyp_pad, device=dev) for _ in range(world)]
        dist.all_gather(gathered, hyp_pad.to(dev))
        gathered = [g.cpu() for g in gathered]
    else:
        gathered = [hyp_pad]
    aggregate_check = None
    if rank == 0:
        from oracle import per_oracle as po
        hyps_all = [None] * n_total
        for r in range(world):
            for j, i in enumerate(parallel.shard_indices(n_total, r, world)):
                row = gathered[r][j]
                hyps_all[i] = row[row >= 0].numpy().astype(np.int32)
        d = po.levenshtein_batch(refs_all, hyps_all)
        per_single = [po.per_from_counts(int(x), len(r), len(h)) for x, r, h in zip(d, refs_all, hyps_all)]
        same = (res_dev["per_scores"] == per_single and res_dev["per"] == np.mean(per_single)
                and res_dev["per_std"] == np.std(per_single) and res_e2e["per_scores"] == per_single)
        aggregate_check = {"utterances": n_total, "sharded_equals_single_process": bool(same),
                           "checker": "CPU oracle (oracle/per_oracle.c) over every rank's hypotheses"}
        assert same, "the sharded PER aggregate differs from the single-process aggregate over the same utterances"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- where the pass goes: each phase timed on its own (CUDA events, device-resident inputs) --------------------
    from whisper_ipa_b200.audio import log_mel_features
    rf, ro = metrics._pack(refs_all[:B])
    rf_d, ro_d = torch.from_numpy(rf).to(dev), torch.from_numpy(ro).to(dev)
    max_ref = int(np.max(np.diff(ro)))

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out

    t_mel, mel_dev = timed(lambda: log_mel_features(audio_dev, arch.n_mels))
    t_enc, _ = timed(lambda: model.encoder(mel_dev, return_features=False))
    t_dec, (ids_d, lens_d) = timed(lambda: model.decode_tokens(tr.prompt, args.max_new, num_beams=args.beams))
    t_per, _ = timed(lambda: tr.score_device(ids_d, lens_d, rf_d, ro_d, max_ref))
    steps_per_pass = P - 1 + args.max_new
    phases = {"logmel_ms": t_mel, "encoder_ms": t_enc, "decode_ms": t_dec, "per_ms": t_per,
              "decode_us_per_step": 1000.0 * t_dec / steps_per_pass}
    del mel_dev

    # ---- roofline of the dominant kernel: the cross-attention streamer timed alone ---------------------------------
    esz = 4 if args.dtype == "float32" else 2
    h16 = "bf16" if args.dtype == "bfloat16" else "f16"
    tdt = _lib.torch_h16(h16)
    st = torch.cuda.current_stream().cuda_stream
    lib = model._lib
    reps = 5
    S = B * args.beams
    H, dm, L, ffn, V = arch.heads, arch.d_model, arch.dec_layers, arch.ffn, arch.vocab
    latent = bool(model.info().get("xattn_latent", 0))
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    shape_key = f"{args.arch}/B{B}/beams{args.beams}/{h16 if esz == 2 else 'f32'}"
    if latent:
        # latent cross-attention (attn_lat.cu): one pass over the encoder output E [B, 1500, d] per layer serves every head and
        # both the key and the value role.  Timed on its own buffers of the decode shapes; E (0.59 GB at 256 clips) is far
        # larger than L2, so every launch re-streams it from HBM.
        tiled = bool(int(os.environ.get("WIPA_XL_TILED", "1")))          # the layout the context keeps (default: chunk-tiled)
        E = torch.randn(B, 1500, dm, device=dev).to(tdt)
        if tiled:
            Et = torch.zeros(B * lib.wipa_test_lat_tiled_elems(H, 1500), device=dev, dtype=tdt)
            _lib.check(lib.wipa_test_lat_tile(E.data_ptr(), B, 1500, H, Et.data_ptr(), st), "lat_tile", lib)
            E = Et
        Qp = (torch.randn(S, H, dm, device=dev) * (1.5 / dm ** 0.5)).to(tdt)
        Cl = torch.empty(S, H, dm, device=dev, dtype=tdt)
        utt = (torch.arange(S, device=dev, dtype=torch.int32) // args.beams).to(torch.int32)
        n_launch = reps * L

        def launch():
            return lib.wipa_test_cross_attn_latent(Qp.data_ptr(), E.data_ptr(), B, utt.data_ptr(), Cl.data_ptr(), S, H, 1500,
                                                   2 if tiled else 0, args.beams, st)
        _lib.check(launch(), "cross_attn_latent", lib)
        torch.cuda.synchronize()
        r0.record()
        for _ in range(n_launch):
            launch()
        r1.record()
        torch.cuda.synchronize()
        kernel_name, source = "cross_attention_latent_kernel", "whisper_ipa_b200/csrc/attn_lat.cu"
        if H == 20 or (H == 16 and tiled and os.environ.get("WIPA_XL_WIDE16", "1") != "0"):
            # whisper-medium / large*: 128 columns per warp (20 heads: two CTAs of 10 heads per key range)
            kernel_name, source = "cross_attention_latent_wide_kernel", "whisper_ipa_b200/csrc/attn_lat_wide.cu"
        bytes_per_launch = B * 1500 * dm * 2 + 2 * S * H * dm * 2       # E once + absorbed queries in + context rows out
        xattn_step_bytes = L * bytes_per_launch
        del E
    else:
        q = torch.randn(S, dm, device=dev)
        for l in range(L):
            _lib.check(lib.wipa_test_cross_attn(model._ctx, S, l, q.data_ptr(), None, st), "cross_attn", lib)
        torch.cuda.synchronize()
        r0.record()
        for _ in range(reps):
            for l in range(L):     # L distinct K/V caches: the working set cycles through >> L2 bytes
                lib.wipa_test_cross_attn(model._ctx, S, l, q.data_ptr(), None, st)
        r1.record()
        torch.cuda.synchronize()
        n_launch = reps * L
        kernel_name, source = "cross_attention_stream_kernel", "whisper_ipa_b200/csrc/attn.cu"
        bytes_per_launch = B * 2 * 1500 * dm * esz                        # K + V of B utterances, one layer
        xattn_step_bytes = L * bytes_per_launch
    us = 1000.0 * r0.elapsed_time(r1) / n_launch
    achieved = bytes_per_launch / (us * 1e-6) / 1e9
    peak, peak_src = measured_peaks()
    traffic, traffic_src = committed_traffic(kernel_name, source, shape_key)
    ca_share = us * 1e-3 * L * steps_per_pass / (ms_total / K)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "us_per_launch": us,
                "bytes_per_launch": bytes_per_launch, "peak_source": peak_src, "share_of_step_est": ca_share}
    if latent:
        # for orientation only: the K/V formulation SURVEY.md 8(d) counts (2 * 1500 * d * 2 bytes per utterance per layer) would
        # have to move twice the bytes in the same time; `achieved` / `frac` above use the bytes THIS kernel's algorithm needs
        kv_bytes = B * 2 * 1500 * dm * 2
        roofline["kv_formulation_bytes_per_launch"] = kv_bytes
        roofline["kv_formulation_equivalent_gbs"] = kv_bytes / (us * 1e-6) / 1e9

    # the whole decode step against the same peak: algorithmic bytes = cross-attention stream + decoder weights (once per step,
    # incl. the folded projections in latent mode) + self-KV read at the mean length + new K/V written
    q2 = bool(int(os.environ.get("WIPA_XL_Q2STEP", "1")))       # two-step absorbed query: Wq + per-head Wk^T instead of the folded [H*d, d]
    o2 = bool(int(os.environ.get("WIPA_XL_O2STEP", "1")))       # two-step context: per-head Wv + Wo instead of the folded [d, H*d]
    w_dec = ((L * (6 * dm * dm + 2 * dm * ffn) + V * dm) * esz if not latent
             else (L * (4 * dm * dm + (2 if q2 else H) * dm * dm + (2 if o2 else H) * dm * dm + 2 * dm * ffn) + V * dm) * esz)
    t_mean = P + (args.max_new - 1) / 2.0
    self_kv = S * L * 2 * t_mean * dm * esz + S * L * 2 * dm * esz
    step_bytes = xattn_step_bytes + w_dec + self_kv
    step_us = phases["decode_us_per_step"]
    roofline["decode_step"] = {"bytes": step_bytes, "us": step_us, "achieved": step_bytes / (step_us * 1e-6) / 1e9,
                               "frac": step_bytes / (step_us * 1e-6) / 1e9 / peak, "unit": "GB/s",
                               "terms": {"cross_attention": xattn_step_bytes, "weights": w_dec, "self_kv": self_kv}}
    if esz == 2 and args.beams == 1:
        # the node that ends the step: vocabulary projection fused with the masked argmax (V * d weights streamed once, logits
        # never stored).  L2 (126 MB) would hold the 80 MB matrix between back-to-back launches, which the real step never
        # allows: flush it before every timed launch and time each launch on its own.
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        _lib.check(lib.wipa_test_logits_argmax(model._ctx, S, st), "logits_argmax", lib)
        ts = []
        for _ in range(8):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.wipa_test_logits_argmax(model._ctx, S, st)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1000.0)
        lus = float(np.median(ts))
        lbytes = V * dm * 2 + S * dm * 2
        roofline["logits_argmax"] = {"bytes": lbytes, "us": lus, "achieved": lbytes / (lus * 1e-6) / 1e9,
                                     "frac": lbytes / (lus * 1e-6) / 1e9 / peak, "unit": "GB/s", "l2": "flushed before each launch"}
        del flush

    parity = None
    if not args.no_parity and args.beams == 1:
        parity = parity_vs_hf_fp32(model, hf_model, audio_b, tr.prompt, args.max_new, res_dev["local_hypotheses"])
    gpu_baseline = None
    if not args.no_gpu_baseline:
        gpu_baseline = hf_gpu_baseline(hf_model, audio_b, tr.prompt, args.max_new, B, beams=args.beams)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        import logging
        import warnings
        from transformers import WhisperFeatureExtractor
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        warnings.filterwarnings("ignore")
        logging.getLogger("transformers").setLevel(logging.ERROR)
        n = args.cpu_baseline_clips
        fe = WhisperFeatureExtractor(feature_size=arch.n_mels)
        dt, _ = cpu_oracle_pass(hf_model, fe, audio_b[:n].numpy(), refs_all[:n], tr.prompt, args.max_new)
        cpu_baseline = {"value": n * CLIP_SECONDS / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"{n} clips x {args.max_new} greedy tokens + PER, HF transformers fp32 (the parity oracle) on {cores} host threads, {dt:.1f} s"}

    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"float16": "f16", "bfloat16": "bf16", "float32": "f32"}[args.dtype], "data": "synthetic",
        "config": {"workload": f"whisper-{args.arch} {'greedy' if args.beams == 1 else f'beam-{args.beams}'} IPA decode + PER, {B} synthetic "
                               f"30 s clips per GPU per step, {args.max_new} new tokens, random-init weights (BASELINE "
                               f"{CONFIG_OF_ARCH[args.arch]}); one evaluate_ids sweep of {n_total} utterances per timed region",
                   "clips_per_gpu": B, "max_new_tokens": args.max_new, "beams": args.beams, "parallelism": f"dp{world}",
                   "sweep_utterances": n_total,
                   "l2_policy": (f"inputs larger than L2: encoder output {B * 1500 * dm * 2 / 1e9:.2f} GB is re-streamed by every decoder layer of "
                                 f"every step, weights every step" if latent else
                                 f"inputs larger than L2: cross-KV {B * 2 * L * 1500 * dm * esz / 1e9:.2f} GB + weights are re-streamed every decode step"),
                   "cross_attention": "latent (encoder output, folded k/v projections)" if latent else "per-layer cross-KV"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "passes_ms": e2e_ms, "reported": "mean of the timed passes"},
        "gpu_launches": int(launches.item()),
        "clocks": clocks,
        "roofline": roofline,
        "phases": phases,
        "parity": parity,
        "cpu_baseline": cpu_baseline,
        "gpu_baseline": gpu_baseline,
        "per_mean": float(res_dev["per"]),
        "aggregate_check": aggregate_check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
