"""SURVEY 8(f) N2 on the GPU: codec-embedding sum for n frames (full-size bf16 tables: 3072 x 2048 + 15 x 2048 x 2048).
Device-resident timing with CUDA events through q3tts_codec_embed_sum_device, host-to-host timing through
q3tts_codec_embed_sum, and the CPU oracle (torch, all host threads) on the same codes.
usage: python tests/tools/embed_bench.py [n_frames] [iters]"""
import ctypes as C, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np
import torch
import qwen3tts_cuda as q
from oracle import codec_embed as oe
from tools.synth_checkpoint import write_codec_embeddings

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 375
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
with tempfile.TemporaryDirectory() as d:
    write_codec_embeddings(d, dtype="bfloat16", seed=3)
    tables = oe.load_tables(d)
    emb = q.CodecEmbedder(d)
rng = np.random.default_rng(0)
codes = np.stack([rng.integers(0, v, size=n) for v in emb.vocab], axis=1).astype(np.int32)
out = emb(codes)                                                   # warm-up + result
ref = oe.codec_embed_sum(tables, codes)
got = torch.from_numpy(out.astype(np.int32) << 16).view(torch.float32).to(torch.bfloat16)
assert torch.equal(got, ref), "CUDA result differs from the oracle"
d_codes = torch.from_numpy(codes).cuda()
d_out = torch.empty((n, emb.hidden), dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
stream = torch.cuda.current_stream()
L = q.lib()
def run():
    rc = L.q3tts_codec_embed_sum_device(emb._h, d_codes.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
    assert rc == 0
for _ in range(3): run()
ms = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
assert torch.equal(d_out.cpu(), ref)
t_dev = float(np.median(ms))
t0 = time.perf_counter()
for _ in range(5): emb(codes)
t_host = (time.perf_counter() - t0) / 5 * 1e3
t0 = time.perf_counter()
for _ in range(3): oe.codec_embed_sum(tables, codes)
t_cpu = (time.perf_counter() - t0) / 3 * 1e3
bytes_alg = n * emb.hidden * 2 * (emb.groups + 1)                  # 16 rows read + 1 row written per frame
print(json.dumps({"workload": f"codec-embedding sum, {n} frames x {emb.groups} groups, H={emb.hidden}, bf16", "ms_device": round(t_dev, 4),
                  "frames_per_s_device": round(n / t_dev * 1e3), "GBps_algorithmic": round(bytes_alg / t_dev / 1e6, 1),
                  "ms_host_to_host": round(t_host, 3), "frames_per_s_host_to_host": round(n / t_host * 1e3),
                  "ms_cpu_oracle": round(t_cpu, 2), "cpu_threads": torch.get_num_threads(), "bit_exact_vs_oracle": True,
                  "l2": "256 MB flush between timed launches"}))
