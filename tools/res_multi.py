"""Several configurations of the fused residual unit in ONE process (library load and torch import cost 20 s per process on a fresh box).
usage: res_multi.py [rows] [B] [iters]   -- checks small ragged cases first, then times dil 1/3/9 with and without the consumer's snake"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 720000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for dil in (1, 3, 9):
    for osn in (0, 1):
        worst = 0.0
        for b, r in ((3, 20000), (1, 100), (5, 1000)):
            _, d = q.debug_resunit(b, r, dil, osn, q.PREC_FP16, 0)
            worst = max(worst, d)
        ms, d = q.debug_resunit(B, rows, dil, osn, q.PREC_FP16, iters)
        R = B * rows
        print(f"resunit96 dil {dil} out_snake {osn}: small-case max diff {worst:.3e}; rows {R}: {ms:.3f} ms  {2.0*R*96*96*8/ms/1e9:.1f} TF/s  "
              f"{R*96*2*2/ms/1e6:.1f} GB/s  max diff {d:.3e}", flush=True)
