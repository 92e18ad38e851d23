"""GPU debugging aid: per-stage SNR of one precision mode against the fp32 engine / the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
from oracle import decoder as od, weights as ow
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    prec = {"fp16": q.PREC_FP16, "bf16": q.PREC_BF16, "fp32": q.PREC_FP32}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
    B, T = int(sys.argv[3]) if len(sys.argv) > 3 else 1, int(sys.argv[4]) if len(sys.argv) > 4 else 6
    cfg = DecoderConfig.tiny() if which == "tiny" else DecoderConfig()
    st = os.path.join(checkpoint_dir(cfg, seed=7 if which == "tiny" else 20261018), "speech_tokenizer")
    c, w = ow.load_decoder(st)
    codes = synth_codes(cfg, B, T, 1001)
    taps = {}
    ref = od.OracleDecoder(cfg, w, torch.float64).forward(codes, taps).numpy()
    tok = q.Qwen3TTSSpeechTokenizer(st, precision=prec)
    tok.set_taps(True)
    out = tok.decoder(codes)
    for name in ("rvq_sum_first", "rvq_sum_rest", "quantized", "pre_conv", "pre_transformer", "upsample0", "upsample1",
                 "init_conv", "block0", "block1", "block2", "block3", "out_conv"):
        g = tok.stage_tap(name); r = taps[name].numpy()
        print(f"{name:16s} snr {od.snr_db(r, g):7.1f} dB  maxabs {np.abs(g - r).max():.3e}  ref std {r.std():.3f}")
    print("pcm snr", od.snr_db(ref, out), "maxabs", np.abs(out - ref).max())

if __name__ == "__main__":
    main()
