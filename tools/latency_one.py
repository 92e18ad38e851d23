"""Latency of one short decode (BASELINE config 1: B = 1, T = 125, 10 s of audio), host codes in -> host PCM out.
usage: python tools/latency_one.py [iters]      (env Q3TTS_PDL=0 / Q3TTS_GRAPHS=1 select launch variants)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np
import qwen3tts_cuda as q
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes

def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    cfg = DecoderConfig()
    d = os.path.join(checkpoint_dir(cfg), "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(d, precision=q.PREC_FP16)
    codes = np.ascontiguousarray(np.transpose(synth_codes(cfg, 1, 125, 1001), (0, 2, 1)))
    for _ in range(20):
        tok.decode(codes)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        tok.decode(codes)
        ts.append(time.perf_counter() - t0)
    ts = np.sort(np.array(ts)) * 1e3
    print(f"config 1 (B=1, T=125) decode, host to host: p50 {ts[len(ts)//2]:.3f} ms  p10 {ts[len(ts)//10]:.3f}  p90 {ts[9*len(ts)//10]:.3f}  "
          f"PDL={os.environ.get('Q3TTS_PDL', '1')}")
    tok.close()

if __name__ == "__main__":
    main()
