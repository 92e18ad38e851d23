"""One configuration of the GEMM-fused residual unit (conv7 + conv1 in one tcgen05 kernel, block 2: C = 192) for ncu captures.
usage: fuse_one.py C dil with_operand rows B iters    (with_operand: 0 stream only, 1 stream + operand, 2 operand only)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q
C, dil, wo = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rows, B, iters = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
ms, dy, da = q.debug_fused_unit(B, rows, C, dil, wo, q.PREC_FP16, iters)
R = B * rows
print(f"fused unit C {C} dil {dil} operand {wo} rows {R}: {ms:.3f} ms  {2.0*R*C*C*8/ms/1e9:.1f} TF/s  max diff y {dy:.3e} a {da:.3e}", flush=True)
