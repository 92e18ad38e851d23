#!/usr/bin/env python
"""bench.py -- decoded audio-seconds per second of the speech-tokenizer decoder (codes -> 24 kHz PCM).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16|bf16|fp32]
                  [--workload config2|config3|config4|config5]

Workloads (BASELINE.json configs[1..4]):
  config2 (default)  one step = one decode of 64 utterances x 30 s ([64,16,375] codes) PER GPU: weak scaling of independent
                     batches (utterances never talk to each other, so there is no data-path collective).
  config3            one step = 512 mixed-length utterances (T ~ U{25..750} frames, seed 1003) LPT-sharded by utterance over
                     the N ranks with q3tts_partition_lpt: STRONG scaling, value = all 512 utterances / max-over-ranks time.
  config4            the "lite" checkpoint (fp16 on disk, no encoder): 256 utterances x 30 s split over the N ranks (strong).
  config5            chunked streaming: 128 streams per GPU, 0.5 s chunks (6,6,6,7 frames), causal sliding-window state carry;
                     one step = one batched push of every stream; p50 / p99 push latency (host codes in -> host PCM out).
`value` is measured with the codes already in HBM and the PCM left in HBM (CUDA events on the launch stream, max over ranks);
`e2e` goes through the host-buffer entry points (q3tts_decode / q3tts_decode_varlen / q3tts_stream_push_batch) with pinned host
buffers, copies inside the timed region.  The default run also carries short config-3 and config-5 measurements in
`other_configs`, so the driver's 1/2/4/8-GPU sweep records the sharded mixed-length workload without extra flags.
`--impl reference` times the CPU oracle (torch-CPU restatement of the reference decoder; the Swift/MLX reference cannot be
built in this image) on a bounded sample of the same workload: one full utterance of it per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "decoded audio-sec/sec (codes->24kHz PCM)"
UNIT = "audio-s/s"
B_PER_GPU, T_FRAMES, SEED = 64, 375, 1002          # BASELINE.json configs[1]
C3_UTTS, C3_SEED = 512, 1003                       # configs[2]
C4_BATCH, C4_SEED = 256, 1004                      # configs[3]
C5_STREAMS, C5_CHUNKS, C5_SEED = 128, 40, 1005     # configs[4]: per GPU
SEC_PER_FRAME = 0.08


def flop_per_frame(T):                             # 4.963 GFLOP + the attention term
    return 4.963e9 + 16.3e6 * (T / 500.0)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p["bf16_tflops_sustained"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU legs (the oracle)
def oracle_decoder(st_dir, attn_mode="reference"):
    from oracle import decoder as od, weights as ow    # CPU baseline legs only
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, w = ow.load_decoder(st_dir)
    return od.OracleDecoder(cfg.decoder_config, w, torch.float32, attn_mode=attn_mode)


def cpu_time(dec, codes_b16t):
    t0 = time.perf_counter()
    dec.forward(codes_b16t)
    return time.perf_counter() - t0


def cpu_baseline_block(dec, cfg, synth_codes):
    """config 1 (B=1, T=125, seed 1001) in full, median of 5 after one warm-up; plus one full utterance of
    config 2 (T=375) when the config-1 rate says it fits the CPU budget.  audio-s/s is intensive: a batch of 64 takes 64 x as long."""
    c1 = synth_codes(cfg, 1, 125, 1001)
    cpu_time(dec, c1)
    runs = sorted(cpu_time(dec, c1) for _ in range(5))
    med = runs[2]
    out = {"value": 10.0 / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
           "sample": "config 1 in full (B=1, 125 frames = 10 s of audio, seed 1001): median of 5 runs after 1 warm-up, torch-CPU fp32 "
                     "restatement of the reference decoder (oracle/decoder.py), all host threads",
           "runs_s": [round(r, 3) for r in runs]}
    if med * 3.2 <= 12.0:
        t = cpu_time(dec, synth_codes(cfg, 1, T_FRAMES, SEED))
        out["config2_one_utterance"] = {"value": T_FRAMES * SEC_PER_FRAME / t, "unit": UNIT, "seconds": round(t, 3),
                                        "sample": "utterance 0 of config 2 (B=1, 375 frames = 30 s), once; the batch of 64 is 64 x this time"}
    return out


def reference_arm_encode(args, W, K):
    """--impl reference --workload encode: the CPU restatement of the reference ENCODER, one 10 s utterance per step."""
    from oracle import encoder as oe
    from tools.fixtures import checkpoint_dir
    from tools.q3cfg import DecoderConfig, EncoderConfig
    from tools.synth_checkpoint import synth_audio
    ec = EncoderConfig()
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=ec), "speech_tokenizer")
    cfg_o, w_o = oe.load_encoder(d)
    torch.set_num_threads(os.cpu_count() or 1)
    orc = oe.OracleEncoder(cfg_o, w_o, torch.float32)
    one = synth_audio(1, 10 * ec.sampling_rate, 2000)
    for _ in range(W):
        orc.encode(one)
    t0 = time.perf_counter()
    for _ in range(K):
        orc.encode(one)
    dt = time.perf_counter() - t0
    val = K * 10.0 / dt
    what = "one utterance of 10 s of 24 kHz audio, B=1, per step"
    return {"impl": "reference", "metric": "encoded audio-seconds per second (speech-tokenizer encoder, audio -> 16 x 12.5 Hz codes)", "value": val,
            "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "speech-tokenizer encoder, 10 s utterances", "sample_per_step": what,
                       "note": "CPU restatement of the reference encoder (torch-CPU fp32 oracle, all host threads); Swift + MLX cannot be built in this image."},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": what},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def reference_arm(args, cfg, W, K, workload_name):
    from tools.fixtures import checkpoint_dir
    from tools.synth_checkpoint import synth_codes
    if workload_name == "encode":
        return reference_arm_encode(args, W, K)
    lite = workload_name == "config4"
    st_dir = os.path.join(checkpoint_dir(cfg, dtype="float16" if lite else "float32"), "speech_tokenizer")
    dec = oracle_decoder(st_dir, "causal_sw" if workload_name == "config5" else "reference")
    # one step = ONE full utterance of the workload (a bounded sample of it); calibrated on config 1 so that the whole run fits
    t125 = cpu_time(dec, synth_codes(cfg, 1, 125, 1001))
    budget = 240.0
    if workload_name == "config3":
        rng = np.random.default_rng(C3_SEED)
        lens = rng.integers(25, 751, size=C3_UTTS)
        T = int(np.median(lens))
        what = f"one utterance of the median length of config 3 ({T} frames = {T * SEC_PER_FRAME:.0f} s), B=1"
        seed = C3_SEED
    elif workload_name == "config5":
        T, seed = 25, C5_SEED
        what = "one stream's 2 s (25 frames, four 0.5 s chunks) decoded one-shot in the causal sliding-window mode, B=1"
    else:
        T, seed = T_FRAMES, (C4_SEED if lite else SEED)
        what = f"utterance 0 of the batch in full (B=1, {T} frames = {T * SEC_PER_FRAME:.0f} s)"
    if t125 * (T / 125.0) * (K + W) > budget and T > 125:
        T = 125
        what = f"the first 125 frames (10 s) of utterance 0, B=1 (a full utterance would exceed the {budget:.0f} s CPU budget for {K}+{W} steps)"
    sample = synth_codes(cfg, 1, T, seed)
    for _ in range(W):
        dec.forward(sample)
    t0 = time.perf_counter()
    for _ in range(K):
        dec.forward(sample)
    dt = time.perf_counter() - t0
    val = K * T * SEC_PER_FRAME / dt
    return {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak" if workload_name in ("config2", "config5") else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(workload_name, args), "sample_per_step": what,
                       "note": "CPU restatement of the reference decoder (torch-CPU fp32 oracle, all host threads); Swift + MLX cannot be "
                               "built in this image.  audio-s/s is intensive, so the sample's rate is the workload's rate on this CPU."},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": what + ", per step"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_text(name, args):
    B, T = args.batch, args.frames
    return {
        "config2": f"batch-{B} decode of {T * SEC_PER_FRAME:.0f} s utterances ([{B},16,{T}] int32 codes -> [{B},{T * 1920}] f32 PCM) per GPU, random-init 12Hz tokenizer",
        "config3": f"batch-{C3_UTTS} mixed-length (2-60 s, T ~ U{{25..750}} frames, seed {C3_SEED}) utterances sharded by utterance (LPT) across the GPUs, random-init 12Hz tokenizer",
        "config4": f"pruned 'lite' speech tokenizer (fp16 on disk, no encoder), batch {C4_BATCH} x 30 s split by utterance across the GPUs",
        "config5": f"chunked streaming decode, {C5_STREAMS} concurrent streams per GPU, 0.5 s chunks (6,6,6,7 frames) with causal conv/attention state carry, {C5_CHUNKS} chunks per stream",
    }[name]


# ---------------------------------------------------------------------------------------------------- GPU side
class Ctx:
    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.W, self.K = max(args.warmup, 0), max(args.steps, 1)

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def allmax(self, v):
        if self.world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return float(v)

    def allsum(self, v):
        if self.world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            return float(t.item())
        return float(v)

    def timed_device(self, step, W, K, sample_clocks=True):
        """W warm-up steps, then exactly K steps between a barrier + synchronize on both sides; CUDA events on the launch stream;
        max over ranks.  Returns (ms for the K steps, clocks)."""
        stream = torch.cuda.current_stream()
        for _ in range(W):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank).start() if sample_clocks else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            step()
        e1.record(stream)
        self.barrier()
        clocks = sampler.stop() if sampler else None
        return self.allmax(e0.elapsed_time(e1)), clocks

    def timed_host(self, step, W, K):
        """Same protocol for host-buffer calls (synchronous: they return when the PCM is in the caller's buffer): wall clock."""
        for _ in range(W):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return self.allmax(dt)


def roofline_block(tok, step, sync, T_for_flops, frames_per_step, ms_per_step, peaks):
    """Roofline of the dominant KERNEL (label = stage.op): every launch of ONE extra profiled step is bracketed by CUDA events on
    the launch stream.  achieved = algorithmic FLOPs (or bytes) of its launches / their summed time."""
    tok.profile_enable(True)
    step()
    sync()
    stages = tok.profile_get()
    kernels = tok.profile_kernels()
    tok.profile_enable(False)
    if not kernels:
        return None
    dom = max(kernels, key=lambda k: k["ms"])
    tot_ms = sum(s["ms"] for s in stages) or 1e-9
    ai = dom["flops"] / max(dom["bytes"], 1.0)
    ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
    if ai >= ridge:
        ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
                "peak_kind": "sustained bf16 (kernel timed inside a long step)"}
    else:
        ach = dom["bytes"] / (dom["ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"]}
    traffic, traffic_src = None, None
    try:   # dram__bytes_read+write per launch of this kernel from the committed `ncu --set full` capture named in the entry
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ent = json.load(f).get(dom["name"])
        if isinstance(ent, dict):
            traffic, traffic_src = ent.get("bytes_per_launch"), {k: ent.get(k) for k in ("capture", "commit", "date", "config")}
    except Exception:
        pass
    roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": dom["name"], "launches": dom["launches"],
                 "ms_per_launch": dom["ms"] / max(dom["launches"], 1),
                 "flops_per_launch": dom["flops"] / max(dom["launches"], 1), "bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
                 "share_of_step": dom["ms"] / tot_ms, "peak_source": peaks["src"],
                 "kernels": [{"name": k["name"], "n": k["launches"], "ms": round(k["ms"], 3),
                              "tflops": round(k["flops"] / max(k["ms"], 1e-9) / 1e9, 1), "gbs": round(k["bytes"] / max(k["ms"], 1e-9) / 1e6, 1)}
                             for k in sorted(kernels, key=lambda k: -k["ms"])[:14]],
                 "stages": [{"name": s["name"], "ms": round(s["ms"], 3), "tflops": round(s["flops"] / max(s["ms"], 1e-9) / 1e9, 1),
                             "gbs": round(s["bytes"] / max(s["ms"], 1e-9) / 1e6, 1)} for s in stages],
                 "whole_step_tflops": frames_per_step * flop_per_frame(T_for_flops) / (ms_per_step / 1e3) / 1e12})
    return roof


def run_config2(cx, q, cfg, st_dir, prec, with_roofline=True):
    from tools.synth_checkpoint import synth_codes
    a = cx.args
    B, T, W, K = a.batch, a.frames, cx.W, cx.K
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, device=cx.local_rank)
    codes = synth_codes(cfg, B, T, SEED + cx.rank)                       # [B,16,T]
    d_codes = torch.from_numpy(codes).cuda()
    d_pcm = torch.empty((B, T * 1920), dtype=torch.float32, device="cuda")
    d_len = torch.empty(B, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        tok.decode_device(d_codes.data_ptr(), B, T, d_pcm.data_ptr(), d_len.data_ptr(), stream.cuda_stream)

    for _ in range(min(W, 1)):
        step()
    torch.cuda.synchronize()
    launches0 = tok.launch_count()
    ms, clocks = cx.timed_device(step, max(W - 1, 0), K)
    tok.sync(stream.cuda_stream)
    launches = (tok.launch_count() - launches0) * K // (K + max(W - 1, 0))
    audio_s = cx.world * B * T * SEC_PER_FRAME
    value = audio_s * K / (ms / 1e3)

    # e2e: host buffers through q3tts_decode (pinned codes in, PCM + lengths out, copies inside the timed region)
    h_codes = torch.from_numpy(np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))).pin_memory()   # [B,T,16]
    h_pcm = torch.empty((B, T * 1920), dtype=torch.float32).pin_memory()
    h_len = torch.empty(B, dtype=torch.int32).pin_memory()
    L = q.lib()

    def e2e_step():
        st = L.q3tts_decode(tok._h, h_codes.data_ptr(), B, T, 1, h_pcm.data_ptr(), h_len.data_ptr())
        if st != 0:
            raise RuntimeError(L.q3tts_last_error().decode())

    e2e_dt = cx.timed_host(e2e_step, 1, K)
    e2e_val = audio_s * K / e2e_dt
    assert int(h_len[0]) == T * 1920 and float(h_pcm.abs().max()) <= 1.0
    assert torch.equal(h_pcm[:2].cuda(), d_pcm[:2])                      # both entry points run the same chain
    roof = roofline_block(tok, step, lambda: tok.sync(stream.cuda_stream), T, B * T, ms / K, load_peaks()) if with_roofline else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": cx.world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[a.precision], "data": "synthetic",
            "config": {"workload": workload_text("config2", a), "l2": "per-step activation working set >> 126 MB L2 (no flush needed)",
                       "parallelism": f"utterance-sharded x{cx.world}, no data-path collective", "seed": SEED},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h_codes.numel() * 4),
                    "d2h_bytes_per_step": int(h_pcm.numel() * 4 + h_len.numel() * 4)},
            "roofline": roof}
    return line, tok, codes


def config3_inputs(cx, q, cfg):
    from tools.synth_checkpoint import synth_codes
    rng = np.random.default_rng(C3_SEED)
    lens = rng.integers(25, 751, size=C3_UTTS)                          # 2 .. 60 s
    part = q.partition_lpt(lens, cx.world)                              # identical on every rank (deterministic)
    mine = [i for i in range(C3_UTTS) if part[i] == cx.rank]
    big = synth_codes(cfg, 1, int(lens[mine].sum()) if len(mine) else 1, C3_SEED + 17 * cx.rank)[0].T   # [sum T, 16]: one draw, sliced
    offs = np.zeros(len(mine) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens[mine])
    return lens, mine, np.ascontiguousarray(big[: int(offs[-1])]), offs


def run_config3(cx, q, cfg, st_dir, prec, W, K, tok=None, with_roofline=False):
    """512 mixed-length utterances, LPT-sharded by utterance over the ranks; STRONG scaling."""
    own = tok is None
    if own:
        tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, device=cx.local_rank)
    lens, mine, packed, offs = config3_inputs(cx, q, cfg)
    frames_mine, frames_all = int(offs[-1]), int(lens.sum())
    up = 1920
    d_codes = torch.from_numpy(packed).cuda()
    d_pcm = torch.empty(frames_mine * up, dtype=torch.float32, device="cuda")
    d_len = torch.empty(len(mine), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        tok.decode_varlen_device(d_codes.data_ptr(), offs, d_pcm.data_ptr(), d_len.data_ptr(), stream.cuda_stream)

    step()
    torch.cuda.synchronize()
    launches0 = tok.launch_count()
    step()
    launches = tok.launch_count() - launches0
    ms, clocks = cx.timed_device(step, max(W - 2, 0), K)
    tok.sync(stream.cuda_stream)
    audio_s = frames_all * SEC_PER_FRAME
    value = audio_s * K / (ms / 1e3)
    h_codes = torch.from_numpy(packed).pin_memory()
    h_pcm = torch.empty(frames_mine * up, dtype=torch.float32).pin_memory()
    h_len = torch.empty(len(mine), dtype=torch.int32).pin_memory()
    L = q.lib()

    def e2e_step():
        st = L.q3tts_decode_varlen(tok._h, h_codes.data_ptr(), offs.ctypes.data, len(mine), h_pcm.data_ptr(), h_len.data_ptr())
        if st != 0:
            raise RuntimeError(L.q3tts_last_error().decode())

    e2e_dt = cx.timed_host(e2e_step, 1, K)
    # each utterance equals its own B = 1 decode: spot-check the shard's shortest and longest
    worst = 0.0
    if len(mine):
        for k in (int(np.argmin(lens[mine])), int(np.argmax(lens[mine]))):
            single, _ = tok.decode(packed[offs[k]:offs[k + 1]][None])
            worst = max(worst, float(np.abs(single[0] - h_pcm.numpy()[offs[k] * up: offs[k + 1] * up]).max()))
    imbalance = cx.allmax(frames_mine) / (frames_all / cx.world)
    roof = None
    if with_roofline:
        roof = roofline_block(tok, step, lambda: tok.sync(stream.cuda_stream), float(np.mean(lens)), frames_mine, ms / K, load_peaks())
    out = {"value": value, "unit": UNIT, "ms_per_step": ms / K, "scaling": "strong", "steps": K,
           "e2e": {"value": audio_s * K / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(cx.allsum(h_codes.numel() * 4)),
                   "d2h_bytes_per_step": int(cx.allsum(h_pcm.numel() * 4 + h_len.numel() * 4))},
           "utterances": C3_UTTS, "frames": frames_all, "audio_s": audio_s, "load_imbalance_max_over_mean": imbalance,
           "max_abs_vs_own_single_decode": cx.allmax(worst), "gpu_launches_per_step_rank0": int(launches), "clocks": clocks,
           "roofline": roof}
    if own:
        tok.close()
    return out


def run_config4(cx, q, cfg, prec, W, K):
    """The lite checkpoint: same decoder architecture, fp16 on disk, no encoder; 256 x 30 s split over the ranks."""
    from tools.fixtures import checkpoint_dir
    from tools.synth_checkpoint import synth_codes
    st_dir = os.path.join(checkpoint_dir(cfg, dtype="float16"), "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, device=cx.local_rank)
    T = T_FRAMES
    rows = [b for b in range(C4_BATCH) if b % cx.world == cx.rank]
    B = len(rows)
    codes = synth_codes(cfg, C4_BATCH, T, C4_SEED)[rows]
    d_codes = torch.from_numpy(np.ascontiguousarray(codes)).cuda()
    d_pcm = torch.empty((B, T * 1920), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        tok.decode_device(d_codes.data_ptr(), B, T, d_pcm.data_ptr(), 0, stream.cuda_stream)

    step()
    torch.cuda.synchronize()
    launches0 = tok.launch_count()
    step()
    launches = tok.launch_count() - launches0
    ms, clocks = cx.timed_device(step, max(W - 2, 0), K)
    tok.sync(stream.cuda_stream)
    audio_s = C4_BATCH * T * SEC_PER_FRAME
    h_codes = torch.from_numpy(np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))).pin_memory()
    h_pcm = torch.empty((B, T * 1920), dtype=torch.float32).pin_memory()
    L = q.lib()

    def e2e_step():
        st = L.q3tts_decode(tok._h, h_codes.data_ptr(), B, T, 1, h_pcm.data_ptr(), None)
        if st != 0:
            raise RuntimeError(L.q3tts_last_error().decode())

    e2e_dt = cx.timed_host(e2e_step, 1, K)
    out = {"value": audio_s * K / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / K, "scaling": "strong", "steps": K,
           "e2e": {"value": audio_s * K / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(cx.allsum(h_codes.numel() * 4)),
                   "d2h_bytes_per_step": int(cx.allsum(h_pcm.numel() * 4))},
           "utterances_per_gpu": B, "checkpoint": "fp16 on disk, no encoder.* tensors, no encoder_config",
           "gpu_launches_per_step_rank0": int(launches), "clocks": clocks}
    tok.close()
    return out


ENC_FLOP_PER_AUDIO_S = None


def encoder_flops_per_audio_second(ec):
    """Algorithmic FLOPs (2 x MAC) of Qwen3TTSSpeechTokenizerEncoder.encode per second of 24 kHz audio (STE.swift:396-443, 545-591,
    684-705, 816-829), attention excluded (it depends on the utterance length)."""
    rate = float(ec.sampling_rate)
    nf, fl = ec.num_filters, 0.0
    fl += rate * nf * ec.kernel_size * 2
    mult = 1
    for r in reversed(ec.upsampling_ratios):
        dim, hid = mult * nf, mult * nf // ec.compress
        fl += rate * (dim * hid * ec.residual_kernel_size + hid * dim) * 2
        rate /= r
        fl += rate * (2 * dim) * dim * (2 * r) * 2
        mult *= 2
    H, I = ec.hidden_size, ec.intermediate_size
    fl += rate * H * (mult * nf) * ec.last_kernel_size * 2
    fl += rate * ec.num_hidden_layers * (4 * H * H + 2 * H * I) * 2
    ds = ec.downsample_stride
    rate /= ds
    fl += rate * H * H * (2 * ds) * 2
    fl += rate * (2 * ec.codebook_dim * H + 16 * ec.codebook_size * ec.codebook_dim) * 2
    return fl


def run_encode(cx, q, W, K, B=64, seconds=10.0, cpu_baseline=True):
    """the speech-tokenizer ENCODER (audio -> codes).  One step = one q3tts_encode of B utterances of `seconds`
    seconds per GPU, host audio in -> host codes out (the only form of the call: value == e2e).  Every rank encodes its own batch."""
    from tools.fixtures import checkpoint_dir
    from tools.q3cfg import DecoderConfig, EncoderConfig
    from tools.synth_checkpoint import synth_audio
    ec = EncoderConfig()
    if cx.rank == 0:
        checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=ec)
    cx.barrier()
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=ec), "speech_tokenizer")
    enc = q.Qwen3TTSSpeechTokenizerEncoder(d, device=cx.local_rank)      # default engine: tensor cores, split fp16 operands
    samples = int(seconds * ec.sampling_rate)
    audio = synth_audio(B, samples, 2000 + cx.rank)
    out = {}

    def step():
        out["codes"] = enc.encode(audio)

    dt = cx.timed_host(step, W, K)
    launches = enc.launch_count()
    audio_s = B * seconds * cx.world
    fl = encoder_flops_per_audio_second(ec)
    res = {"metric": "encoded audio-seconds per second (speech-tokenizer encoder, audio -> 16 x 12.5 Hz codes)", "value": audio_s * K / dt,
           "unit": UNIT, "ms_per_step": dt / K * 1e3, "steps": K, "scaling": "weak", "dtype": "f32 (as split fp16 pairs on the tensor cores)",
           "workload": f"{B} utterances x {seconds:.0f} s of 24 kHz audio per GPU, host audio in -> host codes out, synthetic weights of the default "
                       f"encoder architecture ({enc.num_parameters / 1e6:.1f} M parameters read by encode)",
           "e2e": {"value": audio_s * K / dt, "unit": UNIT, "h2d_bytes_per_step": int(audio.nbytes), "d2h_bytes_per_step": int(out["codes"].nbytes)},
           "gflop_per_audio_s": fl / 1e9, "gpu_launches_per_step": launches,
           "roofline": {"bound": "tensor", "achieved": fl * audio_s / cx.world * K / dt / 1e12, "peak": load_peaks()["tf_sust"], "unit": "TFLOP/s",
                        "frac": fl * audio_s / cx.world * K / dt / 1e12 / load_peaks()["tf_sust"],
                        "peak_kind": "sustained bf16 tensor peak; ALGORITHMIC float32 FLOPs over the whole call, copies included.  The engine executes "
                                     "3 fp16 products per float32 product (split operands, ~22 mantissa bits) plus one element-wise pass per GEMM",
                        "executed_fp16_tflops": 3.0 * fl * audio_s / cx.world * K / dt / 1e12, "traffic": None}}
    if cpu_baseline and cx.rank == 0:
        import torch as _t
        from oracle import encoder as oe
        cfg_o, w_o = oe.load_encoder(d)
        orc = oe.OracleEncoder(cfg_o, w_o, _t.float32)
        _t.set_num_threads(os.cpu_count() or 1)
        one = audio[:1]
        orc.encode(one)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ref = orc.encode(one).numpy()
            ts.append(time.perf_counter() - t0)
        res["cpu_baseline"] = {"value": seconds / float(np.median(ts)), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"one utterance of {seconds:.0f} s, median of 3, torch-CPU fp32 restatement of the reference encoder (oracle/encoder.py)"}
        res["codes_equal_to_cpu_restatement"] = float((ref == out["codes"][:1]).mean())
    enc.close()
    return res


def run_config5(cx, q, cfg, st_dir, prec, n_chunks=C5_CHUNKS, streams_per_gpu=C5_STREAMS):
    """Chunked streaming: one step = one batched push (host codes in -> host PCM out) of every stream of this GPU."""
    from tools.synth_checkpoint import synth_codes
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, attn_mode=q.ATTN_CAUSAL_SW, device=cx.local_rank)
    S = streams_per_gpu
    pattern = (6, 6, 6, 7)
    warm = 4                                                            # allocation, young streams
    sizes = [pattern[i % 4] for i in range(n_chunks + warm)]
    T = sum(sizes)
    codes = synth_codes(cfg, S, T, C5_SEED + cx.rank)                   # [S,16,T]
    frames = [np.ascontiguousarray(codes[s].T) for s in range(S)]
    streams = [tok.open_stream() for _ in range(S)]
    lat, pos = [], 0
    launches0 = tok.launch_count()
    sampler = None
    cx.barrier()
    for i, n in enumerate(sizes):
        chunk = [f[pos:pos + n] for f in frames]
        if i == warm:
            cx.barrier()
            launches0 = tok.launch_count()
            sampler = ClockSampler(cx.local_rank).start()
        t0 = time.perf_counter()
        tok.push_streams(streams, chunk)
        lat.append(time.perf_counter() - t0)
        pos += n
    clocks = sampler.stop() if sampler else None
    steady = np.array(lat[warm:]) * 1e3
    launches = tok.launch_count() - launches0
    audio_s = S * sum(sizes[warm:]) * SEC_PER_FRAME
    total_time = cx.allmax(float(steady.sum()) / 1e3)
    out = {"value": cx.world * audio_s / total_time, "unit": UNIT, "scaling": "weak", "steps": n_chunks,
           "ms_per_step": 1e3 * total_time / n_chunks,
           "chunk_latency_ms": {"p50": cx.allmax(float(np.percentile(steady, 50))), "p99": cx.allmax(float(np.percentile(steady, 99))),
                                "max": cx.allmax(float(steady.max())), "note": "host codes in -> host PCM out per batched push; max over ranks"},
           "streams_per_gpu": S, "streams_total": S * cx.world, "chunk_frames": list(pattern), "context_frames_per_push": 3,
           "realtime_factor": cx.world * audio_s / total_time / (S * cx.world),
           "gpu_launches_per_push": int(launches // n_chunks), "gpu_launches_total": int(launches), "clocks": clocks,
           "e2e": {"value": cx.world * audio_s / total_time, "unit": UNIT,
                   "h2d_bytes_per_step": int(S * cx.world * 6.25 * 16 * 4), "d2h_bytes_per_step": int(S * cx.world * 6.25 * 1920 * 4)}}
    for s in streams:
        s.close()
    tok.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--workload", default=os.environ.get("Q3TTS_BENCH_WORKLOAD", "config2"), choices=["config2", "config3", "config4", "config5", "encode"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short config-3 / config-5 / bf16 side measurements of the default run")
    args = ap.parse_args()
    cx = Ctx(args)
    W = max(cx.W, 3) if args.impl == "ours" else cx.W                   # timing rule: at least 3 warm-up steps
    K = cx.K

    from tools.fixtures import checkpoint_dir
    from tools.q3cfg import DecoderConfig
    from tools.synth_checkpoint import synth_codes
    cfg = DecoderConfig()

    if args.impl == "reference":
        if cx.rank != 0:
            return 0
        print(json.dumps(reference_arm(args, cfg, cx.W, K, args.workload)))
        return 0

    import qwen3tts_cuda as q
    if not torch.cuda.is_available() or q.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible; the CUDA path has no fallback")
    torch.cuda.set_device(cx.local_rank)
    if cx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local_rank))
    prec = {"fp16": q.PREC_FP16, "bf16": q.PREC_BF16, "fp32": q.PREC_FP32}[args.precision]
    if cx.rank == 0:
        checkpoint_dir(cfg)
        if args.workload == "config4":
            checkpoint_dir(cfg, dtype="float16")
    cx.barrier()                                                        # the other ranks find the cached copies
    st_dir = os.path.join(checkpoint_dir(cfg), "speech_tokenizer")
    cx.W = W
    dtype = {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision]

    if args.workload == "config2":
        line, tok, codes = run_config2(cx, q, cfg, st_dir, prec)
        if not args.no_other_configs:
            other = {}
            other["config3"] = run_config3(cx, q, cfg, st_dir, prec, 3, 2, tok=tok)
            tok.close()
            tok = None
            other["config5"] = run_config5(cx, q, cfg, st_dir, prec, n_chunks=24)
            if args.precision == "fp16":                               # the bf16 line BASELINE configs[1] names, beside the fp16 headline
                a2 = argparse.Namespace(**vars(args))
                a2.precision = "bf16"
                cx2 = Ctx(a2)
                cx2.W, cx2.K = 3, min(K, 5)
                l2, t2, _ = run_config2(cx2, q, cfg, st_dir, q.PREC_BF16, with_roofline=False)
                t2.close()
                other["config2_bf16"] = {"value": l2["value"], "unit": UNIT, "ms_per_step": l2["ms_per_step"], "e2e": l2["e2e"], "dtype": "bf16",
                                         "note": "bf16 operands reach 25-27 dB SNR on this decoder (fp16: 44 dB; tests/test_oracle.py, "
                                                 "tests/test_gpu_parity.py), so the headline mode is fp16 at the same tensor-core rate"}
            other["encode"] = run_encode(cx, q, 1, 2, cpu_baseline=cx.world == 1 and not args.no_cpu_baseline)   # row N3, not a BASELINE config
            line["other_configs"] = other
        if tok is not None:
            tok.close()
        if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(oracle_decoder(st_dir), cfg, synth_codes)
    elif args.workload == "encode":
        r = run_encode(cx, q, W, K, cpu_baseline=cx.world == 1 and not args.no_cpu_baseline)
        line = {"metric": r.pop("metric"), "value": r.pop("value"), "unit": UNIT, "n_gpus": cx.world, "steps": K, "warmup": W,
                "ms_per_step": r.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": r.pop("workload"), "parallelism": f"x{cx.world} independent batches, no collective"},
                "gpu_launches": r.get("gpu_launches_per_step", 0) * K, "e2e": r.pop("e2e"), "roofline": r.pop("roofline")}
        if "cpu_baseline" in r:
            line["cpu_baseline"] = r.pop("cpu_baseline")
        r.pop("steps", None); r.pop("scaling", None); r.pop("dtype", None)
        line.update(r)
    else:
        if args.workload == "config3":
            r = run_config3(cx, q, cfg, st_dir, prec, W, K, with_roofline=True)
        elif args.workload == "config4":
            r = run_config4(cx, q, cfg, prec, W, K)
        else:
            r = run_config5(cx, q, cfg, st_dir, prec)
        line = {"metric": METRIC, "value": r.pop("value"), "unit": UNIT, "n_gpus": cx.world, "steps": r.pop("steps"), "warmup": W,
                "ms_per_step": r.pop("ms_per_step"), "higher_is_better": True, "scaling": r.pop("scaling"), "vs_baseline": None,
                "dtype": dtype, "data": "synthetic",
                "config": {"workload": workload_text(args.workload, args), "parallelism": f"utterance-sharded x{cx.world}, no data-path collective",
                           "l2": "per-step activation working set >> 126 MB L2 (no flush needed)"},
                "clocks": r.pop("clocks", None),
                "gpu_launches": r.get("gpu_launches_total", r.get("gpu_launches_per_step_rank0", 0) * K),
                "e2e": r.pop("e2e"), "roofline": r.pop("roofline", None)}
        line.update(r)
        if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
            lite = args.workload == "config4"
            cst = os.path.join(checkpoint_dir(cfg, dtype="float16" if lite else "float32"), "speech_tokenizer")
            line["cpu_baseline"] = cpu_baseline_block(oracle_decoder(cst), cfg, synth_codes)
    if cx.rank == 0:
        print(json.dumps(line))
    if cx.world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
