"""Timing of one GEMM-fused residual unit shape. usage: fuse_one.py C dil with_operand rows B iters"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
C, dil, wo, rows, B, iters = (int(a) for a in sys.argv[1:7])
ms, dy, da = q.debug_fused_unit(B, rows, C, dil, wo, q.PREC_FP16, iters)
R = B * rows
print(f"{os.environ.get('TAG','')} fused unit C {C} dil {dil} operand {wo} rows {R}: {ms:.3f} ms  {2.0*R*C*C*8/ms/1e9:.1f} TF/s  diff y {dy:.3e} a {da:.3e}", flush=True)
