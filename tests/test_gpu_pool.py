"""The in-library multi-GPU scheduler (q3tts_pool_*, north_star (d)): one worker thread + context + model replica per
GPU, utterances LPT-sharded, no collective.  Oracle: every utterance of a pool decode is bit-identical to the same utterance decoded
through q3tts_decode_varlen on one GPU (and that path is pinned to the CPU oracle in test_gpu_parity.py / test_gpu_bench_scale.py).
Runs with however many GPUs the box has (1 is a valid pool)."""
import numpy as np
import pytest
import torch

import qwen3tts_cuda as q
from oracle import decoder as od
from tools.synth_checkpoint import synth_codes

pytestmark = pytest.mark.gpu


def _utterances(cfg, lens, seed):
    return [np.ascontiguousarray(synth_codes(cfg, 1, max(L, 1), seed + i, zero_frac=0.1)[0].T)[:L] for i, L in enumerate(lens)]


def test_pool_equals_single_gpu_varlen_tiny(tiny_dir, tiny_oracle):
    cfg, w, dec = tiny_oracle
    lens = [9, 1, 30, 0, 17, 5, 5, 22, 3, 13, 40, 2]
    utts = _utterances(cfg, lens, 300)
    one = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, device=0)
    want, want_len = one.decode_varlen(utts)
    pool = q.Qwen3TTSDecoderPool(tiny_dir, precision=q.PREC_FP32)
    assert pool.size == q.device_count()
    for _ in range(2):                                           # the second call reuses every buffer
        got, got_len = pool.decode_varlen(utts)
        assert np.array_equal(got_len, want_len)
        for i in range(len(lens)):
            assert np.array_equal(got[i], want[i]), i
    stats = pool.last_stats()
    assert len(stats) == pool.size and sum(s["frames"] for s in stats) == sum(lens)
    # against the oracle too (fp32 tolerance), for one utterance per worker share
    for i in (2, 10):
        ref, ref_len = dec.decode(utts[i][None])
        assert np.abs(got[i] - ref.numpy()[0]).max() <= 1e-4 and got_len[i] == ref_len[0]
    # int16 and the uniform-batch entry point
    got16, _ = pool.decode_varlen(utts, int16=True)
    for f, i16 in zip(want, got16):
        assert np.array_equal(i16, q.pcm_to_int16(f))
    codes = np.ascontiguousarray(np.transpose(synth_codes(cfg, 7, 11, 5), (0, 2, 1)))
    a1, l1 = one.decode(codes)
    a2, l2 = pool.decode(codes)
    assert np.array_equal(a1, a2) and np.array_equal(l1, l2)
    # pinned destination: the workers DMA straight into the caller's buffer
    total = sum(lens) * cfg.total_upsample
    pinned = torch.empty(total, dtype=torch.float32).pin_memory()
    got_p, _ = pool.decode_varlen(utts, out=pinned.numpy())
    for i in range(len(lens)):
        assert np.array_equal(got_p[i], want[i]), i
    # errors surface with the worker's message; the pool stays usable
    bad = [u.copy() for u in utts]
    bad[4][3, 2] = cfg.codebook_size
    with pytest.raises(q.AudioDecodingFailed) as e:
        pool.decode_varlen(bad)
    assert e.value.status == 1 and "GPU" in str(e.value)
    got, _ = pool.decode_varlen(utts)
    assert np.array_equal(got[7], want[7])
    pool.close()
    one.close()


def test_pool_full_model_mixed_lengths_fp16(full_dir, full_oracle):
    # a slice of BASELINE config 3 (mixed 2-60 s utterances, seed 1003): pool == single GPU, bit for bit; one utterance vs the oracle
    cfg, w, dec32 = full_oracle
    rng = np.random.default_rng(1003)
    lens = rng.integers(25, 751, size=24).tolist()
    utts = _utterances(cfg, lens, 1003)
    one = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16, device=0, max_frames_per_launch=3000)   # several micro-batches
    want, want_len = one.decode_varlen(utts)
    pool = q.Qwen3TTSDecoderPool(full_dir, precision=q.PREC_FP16, max_frames_per_launch=3000)
    got, got_len = pool.decode_varlen(utts)
    assert np.array_equal(got_len, want_len)
    for i in range(len(lens)):
        assert np.array_equal(got[i], want[i]), i
    i = int(np.argmin(lens))
    ref, _ = dec32.decode(utts[i][None])
    assert od.snr_db(ref.numpy()[0], got[i]) >= 40.0
    pool.close()
    one.close()
