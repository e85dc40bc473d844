#!/usr/bin/env python
"""bench.py -- codec tokens/s (frames/s) + 44.1 kHz samples/s of the fish-speech hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic utterances per GPU:
prompt (C+1, P) -> prefill -> N frames of dual-AR decode (slow step + 8 fast steps, device-side
sampling) -> Firefly vocoder -> PCM.  Default workload = BASELINE.json configs[4] ("cfg5": Fish 1.5,
long-form 60 s utterances = 1292 frames, P = 384, temp 0.7 / top_p 0.8), 32 utterances per GPU: the
largest single-GPU configuration, and the per-GPU shard of the 8-GPU batch of 256.  With N GPUs the
32 N utterances are assigned to ranks by fish_speech_rs_b200.shard.assign (independent utterances, no
collective on the data path, "weak" scaling).  At N = 1 the line also carries two short secondary
blocks on the same handle: `cfg3` (B = 16 mixed prompts, 216 frames) and `latency_b1` (cfg2, B = 1).

  value  : frames/s from device time only (CUDA events on the library's stream: prefill + frame loop
           + vocoder kernels), inputs resident in HBM
  e2e    : frames/s by wall clock through the C ABI with HOST buffers (pinned), copies included
  roofline: weight-streaming GEMV kernel, algorithmic bytes / CUDA-event duration per launch,
           against MEASURED_PEAKS.json
  cpu_baseline: the oracle (restated Candle-CPU path, PyTorch CPU fp32) timed on a bounded sample

`--impl reference` times the oracle alone on the host cores (the Rust/Candle reference cannot be built
here: no cargo).  Only this file's cpu legs import `oracle/`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: fish version, utterances per GPU, prompt lengths, frames, sampling
    "cfg1": dict(version="1.5", batch=1, prompt_lens=lambda b: [330] * b, frames=40, temp=0.0, top_p=0.8,
                 desc="Fish 1.5 greedy, single short utterance, default voice (P=330, N=40): the reference's CPU plumbing case"),
    "cfg2": dict(version="1.5", batch=1, prompt_lens=lambda b: [384] * b, frames=216, temp=0.7, top_p=0.8,
                 desc="Fish 1.5 B=1 temp=0.7 top_p=0.8 10 s utterance (P=384, N=216)"),
    "cfg3": dict(version="1.5", batch=16, prompt_lens=lambda b: [300 + 28 * (i % 16) for i in range(b)], frames=216,
                 temp=0.7, top_p=0.8, desc="Fish 1.5 B=16 mixed prompts 300..720, N=216 each"),
    "cfg4": dict(version="1.4", batch=1, prompt_lens=lambda b: [274 + 110] * b, frames=216, temp=0.7, top_p=0.8,
                 clip_samples=562265,
                 desc="Fish 1.4 voice clone: 12.75 s clip -> log-mel -> encoder -> 274 code frames -> conditioned decode (N=216) -> vocoder"),
    "b2": dict(version="1.5", batch=2, prompt_lens=lambda b: [384] * b, frames=64, temp=0.7, top_p=0.8,
               desc="tuning probe: B=2, P=384, N=64"),
    "b4": dict(version="1.5", batch=4, prompt_lens=lambda b: [384] * b, frames=64, temp=0.7, top_p=0.8,
               desc="tuning probe: B=4, P=384, N=64"),
    "b8": dict(version="1.5", batch=8, prompt_lens=lambda b: [384] * b, frames=64, temp=0.7, top_p=0.8,
               desc="tuning probe: B=8, P=384, N=64"),
    "cfg5": dict(version="1.5", batch=32, prompt_lens=lambda b: [384] * b, frames=1292, temp=0.7, top_p=0.8,
                 desc="Fish 1.5 long-form 60 s utterances (P=384, N=1292), 32 per GPU (256 on 8 GPUs), sharded data-parallel"),
}
VOCODER_FLOP_PER_FRAME = 2.646e9  # 1323.2 M MAC per code frame
# of which the 90 ResBlock convs: sum over stages of 126 C^2 MAC per step (6 convs x (3 + 7 + 11) taps) x steps per frame
RESBLOCK_FLOP_PER_FRAME = 2.0 * 126 * (256 ** 2 * 32 + 128 ** 2 * 256 + 64 ** 2 * 512 + 32 ** 2 * 1024 + 16 ** 2 * 2048)
# ... plus the 5 upsampling ConvTranspose1d (2 taps per output step): together the convs that run on tcconv_kernel
TC_FLOP_PER_FRAME = RESBLOCK_FLOP_PER_FRAME + 2.0 * 2 * (512 * 256 * 32 + 256 * 128 * 256 + 128 * 64 * 512 + 64 * 32 * 1024 + 32 * 16 * 2048)
FRAME_RATE = 44100.0 / 2048.0  # 21.533 frames/s (reference prints with 21.535, single_batch.rs:292-295)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(cfgname, rank, world=1):
    """This rank's utterances of the global batch (batch-per-GPU x world utterances, seeds by GLOBAL index),
    assigned by the length-balanced static partition of fish_speech_rs_b200.shard."""
    from fish_speech_rs_b200 import shard, synth
    c = CONFIGS[cfgname]
    mcfg = dict(synth.FISH15 if c["version"] == "1.5" else synth.FISH14)
    tok = dict(synth.FISH15_TOKENS if c["version"] == "1.5" else synth.FISH14_TOKENS)
    voice = np.load(os.path.join(ROOT, "tests", "golden", "default_voice.npy"))  # make_prompt stores codes + 1 for <= 1.4
    total = c["batch"] * world
    lens = c["prompt_lens"](total)
    mine = shard.my_shard([P + c["frames"] for P in lens], rank, world)
    prompts = [synth.make_prompt(mcfg, tok, lens[i], seed=1000 + i, voice=voice) for i in mine]
    return c, mcfg, tok, prompts


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_sample(cfgname, lm_w, codec_w, sample_frames, threads):
    """Oracle (restated Candle-CPU path) on a bounded sample of the workload: prefill of one prompt +
    `sample_frames` decode frames + vocoding of those frames; scaled to the full workload's
    prefill : decode : vocode proportions.  Returns (frames/s, description)."""
    import torch
    from oracle import codec as ocodec
    from oracle import dual_ar as olm
    from oracle import generate as ogen
    from oracle import sampling as osamp
    torch.set_num_threads(threads)
    c, mcfg, tok, prompts = build_inputs(cfgname, 0)
    model = olm.DualARTransformer(lm_w, olm.BaseModelArgs(**mcfg), olm.TokenConfig(**tok), c["version"])
    args = osamp.SamplingArgs(c["temp"], c["top_p"], 256, 1.4, seed=1)
    prompt = torch.from_numpy(prompts[0].astype(np.int64))
    with torch.no_grad():
        gen = ogen.SingleBatchGenerator(model, prompt, 100000, args, True, 0, fixed_len=sample_frames,
                                        force_slow=None if c["version"] == "1.5" else [tok["pad_id"]])
        t0 = time.perf_counter()
        frames = [gen.next()]
        t1 = time.perf_counter()
        for _ in range(sample_frames - 1):
            frames.append(gen.next())
        t2 = time.perf_counter()
        codes = torch.tensor(frames, dtype=torch.int64).T[1:].clamp(max=999)[None]
        ocodec.decode(codes, codec_w)
        t3 = time.perf_counter()
    model.clear_slow_layer_caches()
    t_prefill = t1 - t0  # includes the first frame's fast loop
    t_frame = (t2 - t1) / max(sample_frames - 1, 1)
    t_voc = (t3 - t2) / sample_frames
    N, B = c["frames"], c["batch"]
    total = B * (t_prefill + (N - 1) * t_frame + N * t_voc)
    desc = (f"oracle on 1 utterance: prefill P={prompt.shape[1]} ({t_prefill:.2f}s) + {sample_frames - 1} decode frames "
            f"({t_frame * 1e3:.0f} ms/frame) + vocoder on {sample_frames} frames ({t_voc * 1e3:.0f} ms/frame); "
            f"scaled to B={B} x {N} frames (decode cost taken at the prompt's context length: favours the CPU for long utterances)")
    return B * N / total, desc


def make_weights(version, want_lm=True, want_codec=True, with_encoder=False):
    from fish_speech_rs_b200 import synth
    mcfg = synth.FISH15 if version == "1.5" else synth.FISH14
    lm_w = synth.make_lm_weights(mcfg, seed=1234) if want_lm else None
    codec_w = synth.make_codec_weights(seed=4321, with_encoder=with_encoder) if want_codec else None
    return lm_w, codec_w


def frame_weight_bytes_estimate(version, dtype):
    """Weight bytes one frame-step streams (slow stack + constrained head + 8 x (fast stack + fast head)), from the model
    shapes alone -- the GPU arm reports the library's exact figure in `roofline.frame_bytes`; this one only feeds the
    `config.l2` note, which both arms must print identically."""
    from fish_speech_rs_b200 import synth
    m = synth.FISH15 if version == "1.5" else synth.FISH14
    t = synth.FISH15_TOKENS if version == "1.5" else synth.FISH14_TOKENS
    D, I = m["dim"], m["intermediate_size"]
    layer = (m["n_head"] + 2 * m["n_local_heads"]) * m["head_dim"] * D + D * D + 3 * I * D
    end = t.get("semantic_end_id")
    n_slow = (end - t["semantic_start_id"] + 2) if end is not None else 2  # Fish <= 1.4: the two-way PAD / EOS head (Q8)
    elems = m["n_layer"] * layer + n_slow * D + m["num_codebooks"] * (m["n_fast_layer"] * layer + m["codebook_size"] * D)
    return elems * (2 if dtype == "bf16" else 4)


def config_dict(cfgname, c, world, per_gpu, frames, dtype):
    """`config` of the JSON line: the workload, identical for the GPU arm and the reference arm."""
    return {"workload": cfgname, "desc": c["desc"], "utterances_per_gpu": per_gpu, "utterances_total": c["batch"] * world,
            "frames": frames, "sharding": "fish_speech_rs_b200.shard.assign (length-balanced static partition, no collective)",
            "weights": "seeded random init, Fish %s shapes" % c["version"], "codec_dtype": "f32",
            "l2": "no flush: per-frame weight stream (%.1f GB) >> 126 MB L2" % (frame_weight_bytes_estimate(c["version"], dtype) / 1e9)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    c = CONFIGS[a.config]
    threads = os.cpu_count() or 1
    lm_w, codec_w = make_weights(c["version"])
    vals = []
    desc = ""
    for i in range(a.warmup + a.steps):
        v, desc = cpu_reference_sample(a.config, lm_w, codec_w, a.cpu_sample_frames, threads)
        if i >= a.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": "codec_tokens_per_sec", "value": v, "unit": "frames/s", "n_gpus": a.gpus,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": c["batch"] * c["frames"] / v * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_dict(a.config, c, max(a.gpus, 1), c["batch"], a.frames or c["frames"], a.dtype),
           "extrapolated": "value = the bounded sample of cpu_baseline.sample scaled to the workload; ms_per_step is the "
                           "extrapolated time of one whole step, not the time this run took",
           "audio_samples_per_sec": v * 2048, "rtf": v / FRAME_RATE,
           "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": desc},
           "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def fp32_peak():
    """FP32 FMA peak measured on this pool's B200 by tools/fp32_peak.cu (profiles/r02_fp32_peak.json), else nominal."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_fp32_peak.json")))
        return float(d["fp32_tflops"]), "measured (tools/fp32_peak.cu)"
    except Exception:
        return 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"


class Workload:
    """One configuration bound to a (possibly shared) LM + codec handle: `step()` is one pass of the hot path."""

    def __init__(self, name, lm, codec, rank, world, seed):
        import torch
        from fish_speech_rs_b200 import SamplingArgs
        self.name, self.lm, self.codec = name, lm, codec
        self.c, self.mcfg, self.tok, self.prompts = build_inputs(name, rank, world)
        c = self.c
        self.B, self.N = len(self.prompts), c["frames"]
        self.sargs = SamplingArgs(c["temp"], c["top_p"], 256, 1.4, seed=seed)
        self.clip = None
        if "clip_samples" in c:  # cfg4: the voice comes out of the GPU encoder (FireflyCodec::encode)
            rng = np.random.default_rng(17)
            self.clip = (0.1 * rng.standard_normal(c["clip_samples"])).astype(np.float32)
        self.pin_prompts = []
        for p in self.prompts:
            t = torch.empty(p.shape, dtype=torch.int32).pin_memory()
            v = t.numpy().view(np.uint32)
            v[...] = p
            self.pin_prompts.append(v)
        self.pcm_pin = [torch.empty((1, 1, 2048 * self.N), dtype=torch.float32).pin_memory().numpy() for _ in range(self.B)]
        self.h2d = sum(p.nbytes for p in self.prompts) + self.B * 8 * self.N * 4 + (self.clip.nbytes if self.clip is not None else 0)
        self.d2h = self.B * 8 * self.N * 4 + self.B * 2048 * self.N * 4

    def step(self):
        from fish_speech_rs_b200 import generate_static_batch
        import torch
        enc_ms = 0.0
        if self.clip is not None:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            codes = self.codec.encode(self.clip)  # (1, 8, 274); timed: mel front-end + ConvNeXt encoder + FSQ
            enc_ms = (time.perf_counter() - t0) * 1e3
            assert codes.shape[2] == 274
        codes = generate_static_batch(self.lm, self.pin_prompts, 100000, self.sargs, fixed_len=self.N)
        st = self.lm.stats()
        # synthetic-weight LMs emit codes >= 1000 that the FSQ table rejects (Q11): harness-side clamp
        cl = [np.minimum(x, 999) for x in codes]
        self.codec.decode_batch(cl, out=self.pcm_pin)
        cs = self.codec.stats()
        dev_ms = st["prefill_ms"] + st["decode_ms"] + cs["device_ms"] + enc_ms
        return dev_ms, st, cs, sum(x.shape[1] for x in codes), enc_ms


def time_workload(wl, steps, warmup, barrier):
    for _ in range(warmup):
        wl.step()
    barrier()
    t0 = time.perf_counter()
    acc = dict(dev=0.0, frames=0, launches=0, pre=0.0, dec=0.0, voc=0.0, enc=0.0, dom_ms=0.0, dom_n=0, dom_bytes=0)
    for _ in range(steps):
        dev_ms, st, cs, nf, enc = wl.step()
        acc["dev"] += dev_ms
        acc["frames"] += nf
        acc["launches"] += st["kernel_launches"] + cs["kernel_launches"]
        acc["pre"] += st["prefill_ms"]
        acc["dec"] += st["decode_ms"]
        acc["voc"] += cs["device_ms"]
        acc["enc"] += enc
        acc["dom_ms"] += st["dominant_kernel_ms"]
        acc["dom_n"] += st["dominant_kernel_launches"]
        acc["dom_bytes"] += st["dominant_kernel_bytes"]
    barrier()
    acc["wall_ms"] = (time.perf_counter() - t0) * 1e3
    return acc


def secondary_block(wl, steps, warmup, barrier):
    a = time_workload(wl, steps, warmup, barrier)
    return {"workload": wl.name, "desc": wl.c["desc"], "utterances": wl.B, "frames": wl.N, "steps": steps, "warmup": warmup,
            "value": a["frames"] / (a["dev"] / 1e3), "e2e": a["frames"] / (a["wall_ms"] / 1e3), "unit": "frames/s",
            "ms_per_step": a["dev"] / steps,
            "breakdown_ms_per_step": {"lm_prefill": a["pre"] / steps, "lm_decode": a["dec"] / steps, "vocoder": a["voc"] / steps}}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from fish_speech_rs_b200 import DualARTransformer, FireflyCodec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries rank 0's ONE JSON line and nothing else: native libraries that write to fd 1 (NCCL prints its version
    # banner there) are sent to stderr, the JSON line goes out through a private duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = CONFIGS[a.config]
    extras = [] if (world > 1 or not a.extras or a.config != "cfg5" or a.frames) else ["cfg3", "cfg2"]
    lm_w, codec_w = make_weights(c["version"], with_encoder="clip_samples" in c)
    B, N = c["batch"], c["frames"]
    max_len = max(c["prompt_lens"](B * world)) + N + 8
    for x in extras:
        max_len = max(max_len, max(CONFIGS[x]["prompt_lens"](CONFIGS[x]["batch"])) + CONFIGS[x]["frames"] + 8)
    from fish_speech_rs_b200 import synth
    v15 = c["version"] == "1.5"
    lm = DualARTransformer(lm_w, dict(synth.FISH15 if v15 else synth.FISH14),
                           dict(synth.FISH15_TOKENS if v15 else synth.FISH14_TOKENS), fish_version=c["version"],
                           device=local, dtype=a.dtype, max_batch=B, max_seq_len=max_len, decode_mode=a.decode_mode)
    codec = FireflyCodec(codec_w, fish_version=c["version"], device=local, max_frames=N, with_encoder="clip_samples" in c)
    if not (a.cpu_baseline and rank == 0):
        del lm_w
    wl = Workload(a.config, lm, codec, rank, world, seed=1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        wl.step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    acc = time_workload(wl, a.steps, 0, barrier)
    clocks = sampler.stop()
    if world > 1:
        # per-rank breakdown on stderr (the JSON line on stdout stays rank 0's alone)
        print("[bench rank %d/%d] ms/step: prefill %.1f decode %.1f vocoder %.1f device %.1f wall %.1f | sm %s MHz (max %s) %s" % (
            rank, world, acc["pre"] / a.steps, acc["dec"] / a.steps, acc["voc"] / a.steps, acc["dev"] / a.steps,
            acc["wall_ms"] / a.steps, clocks["sm_mhz"], clocks["sm_max_mhz"], ",".join(clocks["reasons"])),
            file=sys.stderr, flush=True)
    wb = lm.stats()["weight_bytes_per_frame"]
    dom_ms, dom_n, dom_bytes = acc["dom_ms"], acc["dom_n"], acc["dom_bytes"]
    if dom_n > 0:
        kname = ("megab_decode_kernel (wide-batch persistent frame loop: TMA weight ring -> tcgen05.mma with the batch rows as "
                 "the N dimension and three bf16 terms stacked along N, fused FFN + deterministic reduce, GQA attention, per-row samplers" if wl.B > 8 else
                 "mega1_decode_kernel (single-row persistent frame loop: TMA weight ring + register-resident activations; GEMV "
                 "phases + GQA attention + samplers" if wl.B == 1 else
                 "mega_decode_kernel (persistent frame loop: weight-streaming GEMV phases + GQA attention + samplers")
        dom_kernel = kname + "; one launch per utterance batch, timed by CUDA events on its stream in the timed region)"
    else:
        # per-op decode path: one extra, untimed, profiled step with per-launch CUDA events around the GEMV
        from fish_speech_rs_b200 import generate_static_batch
        lm.set_profile(True)
        generate_static_batch(lm, wl.pin_prompts, 100000, wl.sargs, fixed_len=min(N, 12))
        pst = lm.stats()
        lm.set_profile(False)
        dom_ms, dom_n, dom_bytes = pst["dominant_kernel_ms"], pst["dominant_kernel_launches"], pst["dominant_kernel_bytes"]
        dom_kernel = "gemv_kernel (weight-streaming GEMV, fused rmsnorm/residual/swiglu; extra profiled step)"

    t = torch.tensor([acc["dev"], acc["wall_ms"]], dtype=torch.float64, device="cuda")
    fr = torch.tensor([float(acc["frames"]), float(acc["launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = t.tolist()
    frames_all, launches_all = fr.tolist()
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        value = frames_all / (dev_ms_max / 1e3)
        e2e = frames_all / (wall_ms_max / 1e3)
        traffic, traffic_note = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02e_megab_traffic.json" if wl.B > 8 else
                                             ("r01_mega1_traffic.json" if wl.B == 1 else "r01_mega_traffic.json"))))
            if dom_n > 0 and a.dtype == "bf16":
                # DRAM bytes per frame from the committed ncu capture x frames in this launch
                traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["frames_in_launch"] * N
                traffic_note = "ncu dram bytes per frame (%s) x %d frames" % (tr["source"], N)
        except Exception:
            pass
        n_l = max(dom_n, 1)
        avg_ms = dom_ms / n_l
        bytes_per_launch = dom_bytes / n_l
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        fpk, fpk_kind = fp32_peak()
        voc_tf = VOCODER_FLOP_PER_FRAME * acc["frames"] / (acc["voc"] / 1e3) / 1e12 if acc["voc"] > 0 else 0.0
        steps = a.steps
        out = {
            "metric": "codec_tokens_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": a.warmup, "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": config_dict(a.config, c, world, wl.B, N, a.dtype),
            "audio_samples_per_sec": value * 2048, "rtf": value / FRAME_RATE,
            "breakdown_ms_per_step": {"lm_prefill": acc["pre"] / steps, "lm_decode": acc["dec"] / steps,
                                      "vocoder": acc["voc"] / steps, "encoder": acc["enc"] / steps},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                    "audio_samples_per_sec": e2e * 2048, "rtf": e2e / FRAME_RATE},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": dom_kernel,
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_note": traffic_note,
                         "launches_timed": int(dom_n),
                         "avg_launch_us": avg_ms * 1e3, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "frame_bytes": wb,
                         "frame_level_frac": (wb * (N - 1) * steps / (acc["dec"] / 1e3) / 1e9) / peaks["hbm_gbs"]},
            # the ResBlock + upsampling convs (TC_FLOP_PER_FRAME of the VOCODER_FLOP_PER_FRAME algorithmic flops) run on the tensor
            # pipe as 3 fp16 products per fp32-accurate MAC (hi*hi + hi*lo + lo*hi): `issued` counts those, against the
            # measured dense bf16/fp16 throughput; `achieved` stays the algorithmic rate (the FP32-FMA peak is what the
            # round-1 kernels were bounded by)
            "vocoder_roofline": {"bound": "tensor", "kernel": "tcconv_kernel (HiFi-GAN ResBlock + upsampling convs: tcgen05 implicit GEMM, "
                                 "fp16 hi+lo split, 3 products per MAC) + conv1d_kernel (conv_pre, quantizer upsampling: FP32 FMA)",
                                 "achieved": voc_tf, "issued": 3.0 * voc_tf * TC_FLOP_PER_FRAME / VOCODER_FLOP_PER_FRAME,
                                 "peak": peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")), "peak_kind": peak_kind,
                                 "unit": "TFLOP/s",
                                 "frac": 3.0 * voc_tf * TC_FLOP_PER_FRAME / VOCODER_FLOP_PER_FRAME
                                 / peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")),
                                 "fp32_fma_peak": fpk, "fp32_fma_peak_kind": fpk_kind, "algorithmic_over_fp32_peak": voc_tf / fpk,
                                 "flop_per_frame": VOCODER_FLOP_PER_FRAME, "tensor_core_flop_per_frame": TC_FLOP_PER_FRAME},
            "clocks": clocks,
        }
        for x in extras:  # short secondary blocks on the same handles (not the headline)
            w2 = Workload(x, lm, codec, 0, 1, seed=1234)
            out["latency_b1" if x == "cfg2" else x] = secondary_block(w2, 2 if x == "cfg2" else 1, 1, barrier)
        if a.cpu_baseline:
            threads = os.cpu_count() or 1
            v, desc = cpu_reference_sample(a.config, lm_w, codec_w, a.cpu_sample_frames, threads)
            out["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": desc}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    lm.close()
    codec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5", choices=sorted(CONFIGS))
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the secondary cfg3 / latency_b1 blocks of the default N=1 run")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-sample-frames", type=int, default=16)
    ap.add_argument("--decode-mode", type=int, default=0, help="0 auto, 1 per-op kernels + CUDA graph, 2 megakernel")
    ap.add_argument("--frames", type=int, default=0, help="override the frame count (profiling runs only)")
    a = ap.parse_args()
    if a.frames:
        for c in CONFIGS.values():
            c["frames"] = a.frames
            c["desc"] += f" [frames overridden to {a.frames}: NOT the benchmark workload]"
    if a.impl == "reference":
        return run_reference(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
