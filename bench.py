#!/usr/bin/env python
"""Headline benchmark: whisper-small greedy IPA decode + PER, audio-seconds transcribed per second (RTFx).

One "step" = one pass of the whole hot path over one batch of synthetic clips per GPU:
log-mel -> encoder -> 220 greedy decode steps (4-token prompt, 224 positions) -> PER counts.
`value`  : device-timed (CUDA events, max over ranks), audio already resident in HBM.
`e2e`    : the same passes through the public API (pipeline.Transcriber.evaluate_local, what evaluate_ids runs per rank) from
           pinned HOST buffers: host->device copies (double-buffered), reference upload, PER gather and result read-back
           inside the timed region (K steps per pass; the faster of two passes).
`roofline`: the dominant kernel (decoder cross-attention, a persistent HBM streamer: by default the latent kernel that reads the
           encoder output once per layer; with WIPA_XATTN_LATENT=0 the stream-K kernel over per-layer K/V caches) timed alone
           with CUDA events on the decode shapes (bytes per launch >> L2), achieved GB/s vs MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the parity oracle (HF transformers Whisper on the host cores, fp32) on a bounded
           sample of the same workload.  The reference's own runtime (mlx_whisper on Apple Metal) cannot run here.
"""
from __future__ import annotations

import argparse
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
PROMPT_LEN = 4


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


def cpu_oracle_pass(hf_model, n_clips, max_new, first=0):
    """The oracle (HF transformers, fp32, all host threads) over n_clips: features + generate + PER. Returns seconds."""
    from oracle import per_oracle as po
    from transformers import WhisperFeatureExtractor
    audio = synthetic_audio(n_clips, first)
    refs = synthetic_references(n_clips)
    t0 = time.perf_counter()
    fe = WhisperFeatureExtractor(feature_size=hf_model.config.num_mel_bins)
    feats = fe(list(audio), sampling_rate=16000, return_tensors="pt").input_features
    prompt = torch.tensor([[50258, 50259, 50359, 50363]] * n_clips)
    with torch.no_grad():
        ids = hf_model.generate(feats, decoder_input_ids=prompt, max_new_tokens=max_new, do_sample=False)
    d = po.levenshtein_batch(refs, [np.asarray(r, np.int32) for r in ids.tolist()])
    per = np.mean([po.per_from_counts(int(x), len(r), ids.shape[1]) for x, r in zip(d, refs)])
    return time.perf_counter() - t0, float(per)


def run_reference(args):
    """--impl reference: the reference's CPU arm = the HF oracle on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import logging
    import warnings
    warnings.filterwarnings("ignore")
    logging.getLogger("transformers").setLevel(logging.ERROR)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hf_model, _ = random_init_state_dict(args.arch)
    n_clips = args.ref_clips
    for _ in range(args.warmup):
        cpu_oracle_pass(hf_model, n_clips, args.max_new)
    t = 0.0
    for s in range(args.steps):
        dt, _ = cpu_oracle_pass(hf_model, n_clips, args.max_new, first=s * n_clips)
        t += dt
    value = n_clips * CLIP_SECONDS * args.steps / t
    sample = f"{n_clips} clips x {args.max_new} greedy tokens per step, HF transformers fp32 on {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper-{args.arch} greedy IPA decode + PER (random-init weights, synthetic 30 s clips)",
                   "clips_per_step": n_clips, "max_new_tokens": args.max_new},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="small")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("WIPA_BENCH_BATCH", "256")),
                    help="clips per GPU per step (the micro-batch of the eval sweep; 16 / 64 / 256 are the named points)")
    ap.add_argument("--dtype", default="float16", choices=["float16", "bfloat16", "float32"],
                    help="float16 (default: libwipa.so, logits within 1e-3 of the fp32 oracle), bfloat16 (libwipa_bf16.so) or float32")
    ap.add_argument("--max-new", type=int, default=220)
    ap.add_argument("--ref-clips", type=int, default=8, help="clips per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-baseline-clips", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    import whisper_ipa_b200 as w
    from whisper_ipa_b200 import _lib, metrics, pipeline

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
    model = w.WhisperIPA(args.arch, dtype=args.dtype, max_batch=B)
    model.load_state_dict(sd)
    del sd
    tr = pipeline.Transcriber(model, max_new=args.max_new)

    # this rank's clips: global clip index = step-independent (same audio every step; weights / audio resident)
    n_total = B * world
    mine = list(range(rank, n_total, world))
    audio_host = torch.from_numpy(synthetic_audio(B, first=rank * B)).pin_memory()
    refs_all = synthetic_references(n_total)
    refs = [refs_all[i] for i in mine]
    rf, ro = metrics._pack(refs)
    rf_d, ro_d = torch.from_numpy(rf).to(dev), torch.from_numpy(ro).to(dev)
    max_ref = int(np.max(np.diff(ro)))
    audio_dev = audio_host.to(dev)
    gathered = [torch.empty((B, 2), dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None

    def device_step():
        ids, lens = tr.transcribe_device(audio_dev)
        counts = tr.score_device(ids, lens, rf_d, ro_d, max_ref)
        if world > 1:
            dist.all_gather(gathered, counts)        # the one collective of the path: 8 bytes per utterance
        return ids, lens, counts

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        device_step()
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
    for _ in range(K):
        ids, lens, counts = device_step()
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
    value = n_total * CLIP_SECONDS * K / (ms_total / 1000.0)

    # ---- e2e: host buffers in, host results out, through the public API -----------------------------------------------
    # Transcriber.evaluate_local over K micro-batches of pinned HOST audio: every step's host->device copy (double-
    # buffered on a side stream by the pipeline), the references' upload, the PER all-gather and the read-back of ids,
    # lengths and counts are inside the timed region.
    audio_all = audio_host.repeat(K, 1).pin_memory() if K > 1 else audio_host
    refs_k = refs * K

    def e2e_pass():
        c, hyps, hyp_lens = tr.evaluate_local(audio_all, refs_k, micro_batch=B)
        if world > 1:
            for k in range(K):
                dist.all_gather(gathered, c[k * B:(k + 1) * B].contiguous())
        return c.cpu(), hyps, hyp_lens

    e2e_pass()                                                  # warm-up: staging buffers, side stream, allocator blocks of K live micro-batches
    sync_all()
    best_ms = None
    for _ in range(2):            # two timed passes of K steps each, the faster one is reported (host-side hiccups are one-off)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        counts_h, hyps_h, lens_h = e2e_pass()
        t1.record()
        sync_all()
        best_ms = t0.elapsed_time(t1) if best_ms is None else min(best_ms, t0.elapsed_time(t1))
    ms2 = torch.tensor([best_ms], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = n_total * CLIP_SECONDS * K / (float(ms2.item()) / 1000.0)
    h2d = B * N_SAMPLES * 4 + rf.nbytes + ro.nbytes
    d2h = B * args.max_new * 4 + B * 4 + B * 8
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- where the pass goes: each phase timed on its own (CUDA events, device-resident inputs) --------------------
    from whisper_ipa_b200.audio import log_mel_features

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
    t_dec, (ids_d, lens_d) = timed(lambda: model.decode_tokens(tr.prompt, args.max_new))
    t_per, _ = timed(lambda: tr.score_device(ids_d, lens_d, rf_d, ro_d, max_ref))
    phases = {"logmel_ms": t_mel, "encoder_ms": t_enc, "decode_ms": t_dec, "per_ms": t_per,
              "decode_us_per_step": 1000.0 * t_dec / (PROMPT_LEN - 1 + args.max_new)}
    del mel_dev

    # ---- roofline of the dominant kernel: the cross-attention streamer timed alone ---------------------------------
    esz = 4 if args.dtype == "float32" else 2
    h16 = "bf16" if args.dtype == "bfloat16" else "f16"
    tdt = _lib.torch_h16(h16)
    st = torch.cuda.current_stream().cuda_stream
    lib = model._lib
    reps = 5
    latent = bool(model.info().get("xattn_latent", 0))
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if latent:
        # latent cross-attention (attn_lat.cu): one pass over the encoder output E [B, 1500, d] per layer serves every head and
        # both the key and the value role.  Timed on its own buffers of the decode shapes; E (0.59 GB at 256 clips) is far
        # larger than L2, so every launch re-streams it from HBM.
        H, dm = arch.heads, arch.d_model
        E = torch.randn(B, 1500, dm, device=dev).to(tdt)
        Qp = (torch.randn(B, H, dm, device=dev) * (1.5 / dm ** 0.5)).to(tdt)
        Cl = torch.empty(B, H, dm, device=dev, dtype=tdt)
        utt = torch.arange(B, device=dev, dtype=torch.int32)
        n_launch = reps * arch.dec_layers

        def launch():
            return lib.wipa_test_cross_attn_latent(Qp.data_ptr(), E.data_ptr(), B, utt.data_ptr(), Cl.data_ptr(), B, H, 1500, st)
        _lib.check(launch(), "cross_attn_latent")
        torch.cuda.synchronize()
        r0.record()
        for _ in range(n_launch):
            launch()
        r1.record()
        torch.cuda.synchronize()
        kernel_name = "cross_attention_latent_kernel"
        bytes_per_launch = B * 1500 * dm * 2 + 2 * B * H * dm * 2       # E once + absorbed queries in + context rows out
        # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/r01_cross_attention_latent_ncu_full.txt:
        # dram__bytes_read.sum 594.70 MB + dram__bytes_write.sum 7.64 MB at small / B=256); other shapes were not captured
        traffic = 594.702336e6 + 7.644672e6 if (args.arch == "small" and B == 256) else None
        del E
    else:
        q = torch.randn(B, arch.d_model, device=dev)
        for l in range(arch.dec_layers):
            _lib.check(lib.wipa_test_cross_attn(model._ctx, B, l, q.data_ptr(), None, st), "cross_attn")
        torch.cuda.synchronize()
        r0.record()
        for _ in range(reps):
            for l in range(arch.dec_layers):     # 12 distinct K/V caches: the working set cycles through >> L2 bytes
                lib.wipa_test_cross_attn(model._ctx, B, l, q.data_ptr(), None, st)
        r1.record()
        torch.cuda.synchronize()
        n_launch = reps * arch.dec_layers
        kernel_name = "cross_attention_stream_kernel"
        bytes_per_launch = B * 2 * 1500 * arch.d_model * esz              # K + V of B utterances, one layer
        # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/r01_cross_attention_stream_ncu_full.txt:
        # dram__bytes_read.sum 1.180631 GB + dram__bytes_write.sum 4.89 MB at small / bf16 / B=256); other shapes were not captured
        traffic = 1.180631e9 + 4.891392e6 if (args.arch == "small" and esz == 2 and B == 256) else None
    us = 1000.0 * r0.elapsed_time(r1) / n_launch
    achieved = bytes_per_launch / (us * 1e-6) / 1e9
    peak, peak_src = measured_peaks()
    steps_per_pass = PROMPT_LEN - 1 + args.max_new
    ca_share = us * 1e-3 * arch.dec_layers * steps_per_pass / (ms_total / K)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "us_per_launch": us, "bytes_per_launch": bytes_per_launch,
                "peak_source": peak_src, "share_of_step_est": ca_share}
    if latent:
        # for orientation only: the K/V formulation SURVEY.md 8(d) counts (2 * 1500 * d * 2 bytes per utterance per layer) would
        # have to move twice the bytes in the same time; `achieved` / `frac` above use the bytes THIS kernel's algorithm needs
        kv_bytes = B * 2 * 1500 * arch.d_model * 2
        roofline["kv_formulation_bytes_per_launch"] = kv_bytes
        roofline["kv_formulation_equivalent_gbs"] = kv_bytes / (us * 1e-6) / 1e9

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        import logging
        import warnings
        warnings.filterwarnings("ignore")
        logging.getLogger("transformers").setLevel(logging.ERROR)
        n = args.cpu_baseline_clips
        dt, _ = cpu_oracle_pass(hf_model, n, args.max_new)
        cpu_baseline = {"value": n * CLIP_SECONDS / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"{n} clips x {args.max_new} greedy tokens + PER, HF transformers fp32 (the parity oracle) on {cores} host threads, {dt:.1f} s"}

    d = arch.d_model
    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"float16": "f16", "bfloat16": "bf16", "float32": "f32"}[args.dtype], "data": "synthetic",
        "config": {"workload": f"whisper-{args.arch} greedy IPA decode + PER, {B} synthetic 30 s clips per GPU per step, "
                               f"{args.max_new} new tokens, random-init weights (BASELINE configs[2])",
                   "clips_per_gpu": B, "max_new_tokens": args.max_new, "parallelism": f"dp{world}",
                   "l2_policy": (f"inputs larger than L2: encoder output {B * 1500 * d * 2 / 1e9:.2f} GB is re-streamed by every decoder layer of "
                                 f"every step, weights every step" if latent else
                                 f"inputs larger than L2: cross-KV {B * 24 * 1500 * d * esz / 1e9:.2f} GB + weights are re-streamed every decode step"),
                   "cross_attention": "latent (encoder output, folded k/v projections)" if latent else "per-layer cross-KV"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches.item()),
        "clocks": clocks,
        "roofline": roofline,
        "phases": phases,
        "cpu_baseline": cpu_baseline,
        "per_mean": float(np.mean([metrics.per_from_counts(int(c[0]), int(c[1]), int(l)) for c, l in zip(counts_h.tolist(), lens_h)])),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
