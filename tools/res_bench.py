"""GPU check + timing of the fused residual unit kernel. usage: res_bench.py [rows] [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for dil in (1, 3, 9):
    for osn in (0, 1):
        ms, d = q.debug_resunit(B, rows, dil, osn, q.PREC_FP16, iters)
        R = B * rows
        print(f"resunit96 dil {dil} out_snake {osn} rows {R}: {ms:.3f} ms  {2.0*R*96*96*8/ms/1e9:.1f} TF/s  {R*96*2*2/ms/1e6:.1f} GB/s  max diff {d:.4e}", flush=True)
