"""Chunked streaming decode (BASELINE config 5).  The reference has no chunked PCM streaming (SURVEY F2:
Qwen3+Streaming.swift streams token ids and emits one final .audio), so the oracle is CHUNK-INVARIANCE: the concatenated
PCM of any chunking must equal the one-shot decode in the same causal sliding-window attention mode -- which in turn is
pinned to the CPU oracle's `causal_sw` mode."""
import numpy as np
import pytest
import torch

import qwen3tts_cuda as q
from oracle import decoder as od
from tools.synth_checkpoint import synth_codes

pytestmark = pytest.mark.gpu


def _chunks(total, pattern):
    out, i, k = [], 0, 0
    while i < total:
        n = min(pattern[k % len(pattern)], total - i)
        out.append((i, i + n))
        i += n
        k += 1
    return out


@pytest.mark.parametrize("pattern", [(6, 6, 6, 7), (1,), (3, 11, 2), (40,)])
def test_fp32_chunked_equals_one_shot_and_oracle_tiny(tiny_dir, tiny_oracle, pattern):
    cfg, w, _ = tiny_oracle
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, attn_mode=q.ATTN_CAUSAL_SW)
    T = 37
    codes = synth_codes(cfg, 1, T, 4242)                      # [1,16,T]
    one_shot = tok.decoder(codes)[0, 0]
    ref = od.OracleDecoder(cfg, w, torch.float64, attn_mode="causal_sw").forward(codes).numpy()[0, 0]
    assert np.abs(one_shot - ref).max() <= 1e-4
    st = tok.open_stream()
    frames = np.ascontiguousarray(codes[0].T)                 # [T,16]
    pcm = np.concatenate([st.push(frames[a:b]) for a, b in _chunks(T, pattern)])
    assert st.frames == T
    assert pcm.shape == one_shot.shape
    assert np.abs(pcm - one_shot).max() <= 2e-5               # same kernels; only the attention key-tile boundaries move
    assert np.abs(pcm - ref).max() <= 1e-4
    st.close()
    tok.close()


def test_batched_push_equals_single_streams_tiny(tiny_dir, tiny_oracle):
    cfg, _, _ = tiny_oracle
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, attn_mode=q.ATTN_CAUSAL_SW)
    T, S = 26, 5
    codes = synth_codes(cfg, S, T, 99)
    frames = [np.ascontiguousarray(codes[s].T) for s in range(S)]
    want = [tok.decoder(codes[s:s + 1])[0, 0] for s in range(S)]
    streams = [tok.open_stream() for _ in range(S)]
    got = [[] for _ in range(S)]
    # ragged: stream s joins at round s and pushes chunks of different sizes (young and old streams in one launch chain)
    pos = [0] * S
    rnd = 0
    while any(p < T for p in pos):
        active = [s for s in range(S) if s <= rnd and pos[s] < T]
        sizes = [min(T - pos[s], 1 + (s + rnd) % 7) for s in active]
        outs = tok.push_streams([streams[s] for s in active], [frames[s][pos[s]:pos[s] + n] for s, n in zip(active, sizes)])
        for s, n, o in zip(active, sizes, outs):
            got[s].append(o)
            pos[s] += n
        rnd += 1
    for s in range(S):
        assert np.abs(np.concatenate(got[s]) - want[s]).max() <= 2e-5, s
        streams[s].close()
    tok.close()


def test_streaming_needs_causal_mode_and_open_stream(tiny_dir):
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32)      # reference attention: full, bidirectional
    with pytest.raises(q.AudioDecodingFailed) as e:
        tok.open_stream()
    assert e.value.status == 6
    tok.close()


def test_fp16_full_model_chunked_snr(full_dir, full_oracle):
    cfg, w, _ = full_oracle
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16, attn_mode=q.ATTN_CAUSAL_SW)
    T = 50                                                      # 4 s: config 5's 0.5 s chunks (6, 6, 6, 7 frames)
    codes = synth_codes(cfg, 2, T, 1005)
    ref = od.OracleDecoder(cfg, w, torch.float32, attn_mode="causal_sw").forward(codes).numpy()[:, 0]
    streams = [tok.open_stream(), tok.open_stream()]
    frames = [np.ascontiguousarray(codes[s].T) for s in range(2)]
    got = [[], []]
    for a, b in _chunks(T, (6, 6, 6, 7)):
        outs = tok.push_streams(streams, [frames[0][a:b], frames[1][a:b]])
        got[0].append(outs[0]); got[1].append(outs[1])
    pcm = np.stack([np.concatenate(g) for g in got])
    snr = od.snr_db(ref, pcm)
    print(f"chunked fp16 vs causal_sw oracle: SNR {snr:.1f} dB")
    assert snr >= 40.0
    one_shot = tok.decoder(codes)[:, 0]
    # two 16-bit runs with different tile boundaries carry independent rounding noise (each ~44 dB below the signal)
    assert od.snr_db(one_shot, pcm) >= 40.0
    for s in streams:
        s.close()
    tok.close()


def test_bad_code_in_a_batched_push_advances_no_stream(tiny_dir, tiny_oracle):
    # a push commits KV / conv state for every stream of its batch: a bad code id must be rejected BEFORE anything is committed,
    # for the offending stream and for the innocent ones pushed with it
    cfg, _, _ = tiny_oracle
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, attn_mode=q.ATTN_CAUSAL_SW)
    T = 18
    codes = synth_codes(cfg, 2, T, 31)
    frames = [np.ascontiguousarray(codes[s].T) for s in range(2)]
    want = [tok.decoder(codes[s:s + 1])[0, 0] for s in range(2)]
    streams = [tok.open_stream(), tok.open_stream()]
    got = [[], []]
    outs = tok.push_streams(streams, [frames[0][:5], frames[1][:5]])
    got[0].append(outs[0]); got[1].append(outs[1])
    bad = frames[1][5:11].copy()
    bad[3, 2] = cfg.codebook_size                              # acoustic id out of range in the middle of the chunk
    with pytest.raises(q.AudioDecodingFailed) as e:
        tok.push_streams(streams, [frames[0][5:11], bad])
    assert e.value.status == 1
    assert [s.frames for s in streams] == [5, 5]              # nobody moved
    outs = tok.push_streams(streams, [frames[0][5:], frames[1][5:]])   # the same frames again, valid this time
    got[0].append(outs[0]); got[1].append(outs[1])
    for s in range(2):
        assert np.abs(np.concatenate(got[s]) - want[s]).max() <= 2e-5, s
        streams[s].close()
    tok.close()


def test_model_freed_before_its_streams(tiny_dir, tiny_oracle):
    # the Swift wrapper's deinit order is not under the caller's control: a model freed while streams are open stays alive
    # until the last stream closes (q3tts_model_free only marks it)
    cfg, _, _ = tiny_oracle
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, attn_mode=q.ATTN_CAUSAL_SW)
    codes = synth_codes(cfg, 1, 9, 5)
    want = tok.decoder(codes)[0, 0]
    st1, st2 = tok.open_stream(), tok.open_stream()
    tok.close()                                                # q3tts_model_free with two open streams
    pcm = st1.push(np.ascontiguousarray(codes[0].T))           # still works: the model is alive
    assert np.abs(pcm - want).max() <= 2e-5
    st1.close()
    st2.close()                                                # the last close deletes the model
