"""Stage-by-stage comparison of the CUDA encoder with the oracle (debugging aid).  usage: python tests/tools/enc_debug.py [samples] [B] [tiny|full]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np, torch
import qwen3tts_cuda as q
from oracle import encoder as oe
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig, EncoderConfig
from tools.synth_checkpoint import synth_audio

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 49
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ec = EncoderConfig() if (len(sys.argv) > 3 and sys.argv[3] == "full") else EncoderConfig.tiny()
d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=ec), "speech_tokenizer")
cfg, w = oe.load_encoder(d)
orc = oe.OracleEncoder(cfg, w, torch.float64)
enc = q.Qwen3TTSSpeechTokenizerEncoder(d, precision=q.PREC_FP32 if os.environ.get("Q3TTS_ENC_FP32") == "1" else q.PREC_FP16)
a = synth_audio(B, samples, 100 + samples)
taps, margins = {}, []
want = orc.encode(a, taps, margins).numpy()
enc.set_taps(True)
got = enc.encode(a)
for name in ["hid0", "res0", "hid1", "res1", "hid2", "res2", "hid3", "res3", "layer3", "seanet", "transformer", "downsample"]:
    g, wv = enc.stage_tap(name), taps[name].numpy()
    dd = np.abs(g - wv)
    print(f"{name:12s} shape {g.shape} max diff {dd.max():.3e} (scale {np.abs(wv).max():.3f}) worst at {np.unravel_index(dd.argmax(), dd.shape)}")
    if dd.max() > 1e-3 and g.size <= 64:
        print("   got ", np.round(g.ravel(), 4)); print("   want", np.round(wv.ravel(), 4))
print("codes equal:", np.array_equal(want, got), " mismatching entries:", int((want != got).sum()), "of", want.size)
