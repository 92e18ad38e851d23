"""BASELINE config 3: 512 mixed-length utterances (2-60 s) sharded by utterance over the GPUs of one box with the
library's LPT scheduler (q3tts_partition_lpt), no data-path collective.  Every rank decodes its shard with
q3tts_decode_varlen (host buffers) and spot-checks utterances against their own B=1 decode.
usage: [torchrun ...] python tools/config3_bench.py [n_utterances] [seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
import numpy as np
import qwen3tts_cuda as q
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1003
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    cfg = DecoderConfig()
    rng = np.random.default_rng(seed)
    lens = rng.integers(25, 751, size=N)                                      # 2 .. 60 s
    part = q.partition_lpt(lens, world)                                       # identical on every rank (deterministic)
    mine = [i for i in range(N) if part[i] == rank]
    st_dir = os.path.join(checkpoint_dir(cfg), "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(st_dir, precision=q.PREC_FP16, device=local)
    utts = [np.ascontiguousarray(synth_codes(cfg, 1, int(lens[i]), seed + 7 * i)[0].T) for i in mine]
    tok.decode_varlen(utts[:4])                                               # warm-up (workspace allocation)
    t0 = time.perf_counter()
    pcms, lengths = tok.decode_varlen(utts)
    dt = time.perf_counter() - t0
    frames = int(sum(lens[i] for i in mine))
    worst = 0.0
    for k in (0, len(mine) // 2, len(mine) - 1):                              # each utterance == its own B=1 decode
        single, _ = tok.decode(utts[k][None])
        worst = max(worst, float(np.abs(single[0] - pcms[k]).max()))
    print(json.dumps({"config": "3: mixed-length 2-60 s", "rank": rank, "n_gpus": world, "utterances": len(mine), "frames": frames,
                      "load_balance": frames / (float(lens.sum()) / world), "audio_s": frames * 0.08, "seconds": dt,
                      "audio_s_per_s": frames * 0.08 / dt, "max_abs_vs_single_decode": worst}), flush=True)
    tok.close()


if __name__ == "__main__":
    main()
