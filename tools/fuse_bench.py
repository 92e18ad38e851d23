"""GPU check + timing of the GEMM-fused residual unit (C = 192 / 128). usage: fuse_bench.py [rows] [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for C in (192, 128):
    for dil in (1, 9):
        for wo in (1, 0):
            ms, dy, da = q.debug_fused_unit(B, rows, C, dil, wo, q.PREC_FP16, iters)
            R = B * rows
            print(f"fused unit C {C} dil {dil} operand {wo} rows {R}: {ms:.3f} ms  {2.0*R*C*C*8/ms/1e9:.1f} TF/s  {R*C*2*(3+wo)/ms/1e6:.1f} GB/s  diff y {dy:.3e} a {da:.3e}", flush=True)
