"""One configuration of the fused residual unit kernel (for ncu captures). usage: res_one.py dil out_snake [rows] [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
dil, osn = int(sys.argv[1]), int(sys.argv[2])
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 720000
B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ms, d = q.debug_resunit(B, rows, dil, osn, q.PREC_FP16, iters)
R = B * rows
print(f"resunit96 dil {dil} out_snake {osn} rows {R}: {ms:.3f} ms  {2.0*R*96*96*8/ms/1e9:.1f} TF/s  {R*96*2*2/ms/1e6:.1f} GB/s  max diff {d:.4e}", flush=True)
