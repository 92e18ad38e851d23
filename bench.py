#!/usr/bin/env python
"""bench.py -- decoded audio-seconds per second of the speech-tokenizer decoder (codes -> 24 kHz PCM).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16|bf16|fp32]

One "step" = one decode of the workload batch (BASELINE.json configs[1]: 64 utterances x 30 s =
[64,16,375] int32 codes per GPU; weak scaling, utterances are independent so ranks never talk on
the data path).  `value` is measured with the codes already in HBM and the PCM left in HBM
(q3tts_decode_device on torch's current stream, CUDA events on that stream); `e2e` goes through the
host-buffer entry point q3tts_decode with pinned host codes/PCM, copies inside the timed region.
`--impl reference` times the CPU oracle (torch-CPU restatement of the reference decoder; the Swift/MLX
reference cannot be built in this image) on a bounded sample of the same workload.
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
FLOP_PER_FRAME = 4.963e9 + 16.3e6 * (T_FRAMES / 500.0)
CPU_SAMPLE_FRAMES = 25                             # bounded CPU sample: first 2 s of one utterance


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


def oracle_decoder(st_dir):
    from oracle import decoder as od, weights as ow    # CPU baseline legs only
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, w = ow.load_decoder(st_dir)
    return od.OracleDecoder(cfg.decoder_config, w, torch.float32)


def cpu_sample(dec, codes_b16t, reps=1):
    """Times the oracle on codes [1,16,CPU_SAMPLE_FRAMES]; returns audio-s/s."""
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        dec.forward(codes_b16t)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return codes_b16t.shape[0] * codes_b16t.shape[2] * 0.08 / best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 0)
    K = max(args.steps, 1)

    from tools.fixtures import checkpoint_dir
    from tools.q3cfg import DecoderConfig
    from tools.synth_checkpoint import synth_codes
    cfg = DecoderConfig()
    B, T = args.batch, args.frames
    workload = f"batch-{B} decode of {T * 0.08:.0f} s utterances ([{B},16,{T}] int32 codes -> [{B},{T * 1920}] f32 PCM) per GPU, random-init 12Hz tokenizer"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        st_dir = os.path.join(checkpoint_dir(cfg), "speech_tokenizer")
        dec = oracle_decoder(st_dir)
        sample = synth_codes(cfg, 1, T, SEED)[:, :, :CPU_SAMPLE_FRAMES]
        for _ in range(min(W, 1)):
            dec.forward(sample)
        t0 = time.perf_counter()
        for _ in range(K):
            dec.forward(sample)
        dt = time.perf_counter() - t0
        val = K * CPU_SAMPLE_FRAMES * 0.08 / dt
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "note": "CPU restatement of the reference decoder (torch-CPU oracle); Swift+MLX cannot be built here"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"first {CPU_SAMPLE_FRAMES} frames (2 s) of utterance 0, B=1, per step"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA)
    import qwen3tts_cuda as q
    if not torch.cuda.is_available() or q.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible; the CUDA path has no fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prec = {"fp16": q.PREC_FP16, "bf16": q.PREC_BF16, "fp32": q.PREC_FP32}[args.precision]
    model_dir = checkpoint_dir(cfg) if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        model_dir = checkpoint_dir(cfg)     # built by rank 0 above; the others find the cached copy
    st_dir = os.path.join(model_dir, "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, device=local_rank)
    codes = synth_codes(cfg, B, T, SEED + rank)                       # [B,16,T]
    d_codes = torch.from_numpy(codes).cuda()
    d_pcm = torch.empty((B, T * 1920), dtype=torch.float32, device="cuda")
    d_len = torch.empty(B, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        tok.decode_device(d_codes.data_ptr(), B, T, d_pcm.data_ptr(), d_len.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    launches0 = tok.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    tok.sync(stream.cuda_stream)
    ms = e0.elapsed_time(e1)
    launches = tok.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = world * B * T * 0.08
    value = audio_s * K / (ms / 1e3)

    # ---- e2e: host buffers through q3tts_decode (pinned codes in, PCM + lengths out, copies timed)
    h_codes = torch.from_numpy(np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))).pin_memory()   # [B,T,16]
    h_pcm = torch.empty((B, T * 1920), dtype=torch.float32).pin_memory()
    h_len = torch.empty(B, dtype=torch.int32).pin_memory()
    L = q.lib()

    def e2e_step():
        st = L.q3tts_decode(tok._h, h_codes.data_ptr(), B, T, 1, h_pcm.data_ptr(), h_len.data_ptr())
        if st != 0:
            raise RuntimeError(L.q3tts_last_error().decode())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_val = audio_s * K / e2e_dt
    assert int(h_len[0]) == T * 1920 and float(h_pcm.abs().max()) <= 1.0

    # ---- roofline of the dominant KERNEL (label = stage.op), every launch timed live with CUDA events on the launch
    #      stream in one extra profiled step.  achieved = algorithmic FLOPs (or bytes) of its launches / their summed time.
    peaks = load_peaks()
    tok.profile_enable(True)
    step()
    tok.sync(stream.cuda_stream)
    stages = tok.profile_get()
    kernels = tok.profile_kernels()
    tok.profile_enable(False)
    roof = None
    if kernels:
        dom = max(kernels, key=lambda k: k["ms"])
        tot_ms = sum(s["ms"] for s in stages) or 1e-9
        ai = dom["flops"] / max(dom["bytes"], 1.0)
        ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
        per_launch_ms = dom["ms"] / max(dom["launches"], 1)
        if ai >= ridge:
            ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
                    "peak_kind": "sustained bf16 (kernel timed inside a long step)"}
        else:
            ach = dom["bytes"] / (dom["ms"] / 1e3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"]}
        traffic = None
        try:   # dram__bytes_read+write per launch of this kernel from the committed ncu --set full capture (same config)
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom["name"])
        except Exception:
            pass
        roof.update({"traffic": traffic, "kernel": dom["name"], "launches": dom["launches"], "ms_per_launch": per_launch_ms,
                     "flops_per_launch": dom["flops"] / max(dom["launches"], 1), "bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
                     "share_of_step": dom["ms"] / tot_ms, "peak_source": peaks["src"],
                     "kernels": [{"name": k["name"], "n": k["launches"], "ms": round(k["ms"], 3),
                                  "tflops": round(k["flops"] / max(k["ms"], 1e-9) / 1e9, 1), "gbs": round(k["bytes"] / max(k["ms"], 1e-9) / 1e6, 1)}
                                 for k in sorted(kernels, key=lambda k: -k["ms"])[:12]],
                     "stages": [{"name": s["name"], "ms": round(s["ms"], 3), "tflops": round(s["flops"] / max(s["ms"], 1e-9) / 1e9, 1),
                                 "gbs": round(s["bytes"] / max(s["ms"], 1e-9) / 1e6, 1)} for s in stages],
                     "whole_step_tflops": B * T * FLOP_PER_FRAME * K / (ms / 1e3) / 1e12})

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload, "l2": "per-step activation working set >> 126 MB L2 (no flush needed)",
                       "parallelism": f"utterance-sharded x{world}, no data-path collective", "seed": SEED},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h_codes.numel() * 4),
                    "d2h_bytes_per_step": int(h_pcm.numel() * 4 + h_len.numel() * 4)},
            "roofline": roof}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dec = oracle_decoder(st_dir)
        sample = codes[:1, :, :CPU_SAMPLE_FRAMES]
        dec.forward(sample[:, :, :5])
        line["cpu_baseline"] = {"value": cpu_sample(dec, sample), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"first {CPU_SAMPLE_FRAMES} frames (2 s) of utterance 0, B=1, torch-CPU fp32 oracle"}
    if rank == 0:
        print(json.dumps(line))
    tok.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
