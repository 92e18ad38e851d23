"""Throughput of the speech-tokenizer ENCODER (row N3): audio-seconds encoded per second, host audio in -> host codes out.
usage: python tools/encode_bench.py [B] [seconds] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np
import qwen3tts_cuda as q
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig, EncoderConfig
from tools.synth_checkpoint import synth_audio

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=EncoderConfig()), "speech_tokenizer")
    enc = q.Qwen3TTSSpeechTokenizerEncoder(d, precision=q.PREC_FP32 if os.environ.get("Q3TTS_ENC_FP32") == "1" else q.PREC_FP16)
    a = synth_audio(B, int(secs * 24000), 1)
    for _ in range(2):
        codes = enc.encode(a)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter(); enc.encode(a); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    print(f"encode B={B} x {secs:.1f} s: {t*1e3:.2f} ms per call (median of {iters}) = {B*secs/t:.0f} audio-s/s, codes {codes.shape}, "
          f"{enc.num_parameters/1e6:.1f} M parameters")

if __name__ == "__main__":
    main()
