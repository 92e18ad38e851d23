"""BASELINE config 5: chunked streaming decode, N concurrent streams per GPU, 0.5 s chunks (6,6,6,7 frames), causal
sliding-window attention with state carry.  Reports p50 / p99 latency of one batched push (host codes in -> host PCM out)
and aggregate decoded audio-s/s.   usage: python tools/stream_bench.py [streams] [chunks] [precision]
Under torchrun every rank drives its own GPU with `streams` streams (streams are pinned to a GPU for life)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np
import qwen3tts_cuda as q
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    prec = {"fp16": q.PREC_FP16, "bf16": q.PREC_BF16, "fp32": q.PREC_FP32}[sys.argv[3] if len(sys.argv) > 3 else "fp16"]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    cfg = DecoderConfig()
    st_dir = os.path.join(checkpoint_dir(cfg), "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=prec, attn_mode=q.ATTN_CAUSAL_SW, device=local)
    pattern = (6, 6, 6, 7)
    sizes = [pattern[i % 4] for i in range(n_chunks)]
    T = sum(sizes)
    codes = synth_codes(cfg, S, T, 1005 + rank)                              # [S,16,T]
    frames = [np.ascontiguousarray(codes[s].T) for s in range(S)]
    streams = [tok.open_stream() for _ in range(S)]
    lat, pos = [], 0
    for i, n in enumerate(sizes):
        chunk = [f[pos:pos + n] for f in frames]
        t0 = time.perf_counter()
        tok.push_streams(streams, chunk)
        lat.append(time.perf_counter() - t0)
        pos += n
    steady = np.array(lat[4:]) * 1e3                                         # skip the warm-up pushes (allocation, young streams)
    audio_s = S * sum(sizes[4:]) * 0.08
    line = {"workload": f"{S} streams/GPU x {n_chunks} chunks of 0.5 s (6,6,6,7 frames), causal-SW state carry", "n_gpus": world, "rank": rank,
            "chunk_latency_ms": {"p50": float(np.percentile(steady, 50)), "p99": float(np.percentile(steady, 99)), "max": float(steady.max())},
            "audio_s_per_s_per_gpu": audio_s / (steady.sum() / 1e3), "realtime_streams_per_gpu": audio_s / (steady.sum() / 1e3),
            "context_frames_per_push": 3, "launches": tok.launch_count()}
    print(json.dumps(line), flush=True)
    for s in streams:
        s.close()
    tok.close()


if __name__ == "__main__":
    main()
