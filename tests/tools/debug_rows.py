"""GPU debugging aid: where (rows / channels) does a stage differ from the oracle?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
from oracle import decoder as od, weights as ow
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes

which, precs, B, T, stage = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
prec = {"fp16": q.PREC_FP16, "bf16": q.PREC_BF16, "fp32": q.PREC_FP32}[precs]
cfg = DecoderConfig.tiny() if which == "tiny" else DecoderConfig()
st = os.path.join(checkpoint_dir(cfg, seed=7 if which == "tiny" else 20261018), "speech_tokenizer")
c, w = ow.load_decoder(st)
codes = synth_codes(cfg, B, T, 1001)
taps = {}
od.OracleDecoder(cfg, w, torch.float64).forward(codes, taps)
tok = q.Qwen3TTSSpeechTokenizer(st, precision=prec)
tok.set_taps(True)
tok.decoder(codes)
g = tok.stage_tap(stage); r = taps[stage].numpy()
err = np.abs(g - r)            # [B, C, L]
print("shape", g.shape, "nan", np.isnan(g).sum())
per_row = err.max(axis=1)      # [B, L]
per_ch = err.max(axis=2)       # [B, C]
np.set_printoptions(linewidth=200, precision=2, suppress=True)
for b in range(B):
    L = per_row.shape[1]
    bad = np.where(per_row[b] > 0.05)[0]
    print(f"b={b} bad rows {len(bad)}/{L}:", bad[:40], "..." if len(bad) > 40 else "")
    badc = np.where(per_ch[b] > 0.05)[0]
    print(f"b={b} bad channels {len(badc)}/{per_ch.shape[1]}:", badc[:40], "..." if len(badc) > 40 else "")
    if len(bad):
        t = bad[0]
        print("first bad row", t, "got", g[b, :8, t], "want", r[b, :8, t])
