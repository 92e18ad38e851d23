"""GPU micro-benchmark + kernel-level check of the tcgen05 multi-tap GEMM on the decoder blocks' shapes.
usage: python tools/gemm_bench.py [rows_per_utt_at_block3] [B] [iters]   (env Q3TTS_TC_* select kernel variants)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import qwen3tts_cuda as q

def main():
    rows3 = int(sys.argv[1]) if len(sys.argv) > 1 else 720000      # block3 rows per utterance (30 s)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    only = sys.argv[4] if len(sys.argv) > 4 else ""
    # (name, rows divisor from block3 rate, Cin, N, taps, dil, mode)
    shapes = []
    for blk, (c, div, r) in enumerate([(768, 60, 8), (384, 12, 5), (192, 3, 4), (96, 1, 3)]):
        shapes.append((f"b{blk}.convT", div * r, 2 * c, r * c, 2, 1, 2))
        for d in (1, 9):
            shapes.append((f"b{blk}.conv7d{d}", div, c, c, 7, d, 0))
        shapes.append((f"b{blk}.conv1", div, c, c, 1, 1, 1))
    if os.environ.get("GEMM_SHAPES"):   # name:div:cin:n:taps:dil:mode,...
        shapes = [(f[0], *map(int, f[1:])) for f in (x.split(":") for x in os.environ["GEMM_SHAPES"].split(","))]
    for name, div, cin, n, taps, dil, mode in shapes:
        if only and only not in name:
            continue
        rows = rows3 // div
        ms, dy, da = q.debug_conv_gemm(B, rows, cin, n, taps, dil, mode, q.PREC_FP16, iters)
        R = B * rows
        flop = 2.0 * R * taps * n * cin
        io = 2.0 * R * (cin + n * (1 + (mode in (1, 2)) + (mode == 1)))
        print(f"{name:12s} rows {R:9d} K {taps*cin:5d} N {n:5d}  {ms:8.3f} ms  {flop/ms/1e9:7.1f} TF/s  {io/ms/1e6:7.1f} GB/s  "
              f"diff y {dy:.3e} a {da:.3e}", flush=True)

if __name__ == "__main__":
    main()
