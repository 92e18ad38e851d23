"""GPU parity tests of the speech-tokenizer ENCODER: the CUDA path, called through the C ABI
(`q3tts_encode`, ctypes mirror `Qwen3TTSSpeechTokenizerEncoder.encode`), against the CPU oracle on the same synthetic weights
and audio.  Stage outputs: max-abs within 5e-5 of the float64 oracle, relative to the stage's scale.  Codes are integers: they must
be identical except where the oracle's two best distances are within 1e-4 of each other (tools/enc_compare.py)."""
import os

import numpy as np
import pytest
import torch

import qwen3tts_cuda as q
from oracle import encoder as oe
from tools.enc_compare import count_near_tie_frames
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig, EncoderConfig
from tools.synth_checkpoint import synth_audio

pytestmark = pytest.mark.gpu
STAGES = ("hid0", "res0", "hid1", "res1", "hid2", "res2", "hid3", "res3", "layer3", "seanet", "transformer", "downsample")


ENGINES = [pytest.param(q.PREC_FP16, id="tensor_cores"), pytest.param(q.PREC_FP32, id="cuda_cores")]


@pytest.fixture(scope="module", params=ENGINES)
def tiny_pair(request):
    # the tiny architecture's narrow layers fall back to the CUDA-core GEMM inside the tensor-core engine: both routes in one graph
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=EncoderConfig.tiny()), "speech_tokenizer")
    cfg, w = oe.load_encoder(d)
    enc = q.Qwen3TTSSpeechTokenizerEncoder(d, precision=request.param)
    yield d, cfg, oe.OracleEncoder(cfg, w, torch.float64), enc
    enc.close()


@pytest.fixture(scope="module", params=ENGINES)
def full_pair(request):
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=EncoderConfig()), "speech_tokenizer")
    cfg, w = oe.load_encoder(d)
    enc = q.Qwen3TTSSpeechTokenizerEncoder(d, precision=request.param)
    yield d, cfg, oe.OracleEncoder(cfg, w, torch.float64), enc
    enc.close()


def _check(oracle, enc, audio, max_tie_frac=0.03):
    taps, margins = {}, []
    want = oracle.encode(audio, taps, margins).numpy()
    enc.set_taps(True)
    got = enc.encode(audio)
    assert got.dtype == np.int32 and got.shape == want.shape
    for name in STAGES:
        g, w = enc.stage_tap(name), taps[name].numpy()
        assert g.shape == w.shape, (name, g.shape, w.shape)
        scale = max(1.0, float(np.abs(w).max()))
        assert np.abs(g - w).max() <= 5e-5 * scale, (name, float(np.abs(g - w).max()))
    enc.set_taps(False)
    bad = count_near_tie_frames(want, got, [m.numpy() for m in margins], tol=1e-4)
    assert bad <= max(1, int(max_tie_frac * want.shape[0] * want.shape[2])), bad
    return want, got


@pytest.mark.parametrize("B,samples", [(1, 1), (1, 47), (2, 48), (1, 49), (3, 1000), (2, 4801)])
def test_tiny_stages_and_codes_match_the_oracle(tiny_pair, B, samples):
    _, cfg, oracle, enc = tiny_pair
    want, got = _check(oracle, enc, synth_audio(B, samples, 100 + samples))
    assert got.shape == (B, 16, oe.encode_frames(cfg, samples)) == (B, 16, enc.frames(samples))


@pytest.mark.parametrize("B,samples", [(1, 24000), (2, 24000 * 3 + 123), (1, 959), (1, 1921)])
def test_full_architecture_matches_the_oracle(full_pair, B, samples):
    _, cfg, oracle, enc = full_pair
    want, got = _check(oracle, enc, synth_audio(B, samples, 7 + samples))
    assert got.shape[2] == oe.encode_frames(cfg, samples)
    assert enc.hop == 1920 and enc.valid_num_quantizers == 16 and enc.codebook_size == 2048 and enc.sampling_rate == 24000


def test_config1_length_ten_seconds_is_125_frames_and_deterministic(full_pair):
    _, cfg, oracle, enc = full_pair
    a = synth_audio(2, 240000, 1001)
    c1, c2 = enc.encode(a), enc.encode(a)
    assert c1.shape == (2, 16, 125) and np.array_equal(c1, c2)
    assert np.array_equal(enc.encode(a[1:2]), c1[1:2])                 # batch rows are independent
    assert c1.min() >= 0 and c1.max() < 2048
    assert len(np.unique(c1[:, 8:])) > 200                             # the deep codebooks are actually searched
    margins = []
    want = oracle.encode(a, margins=margins).numpy()
    assert count_near_tie_frames(want, c1, [m.numpy() for m in margins], tol=1e-4) <= 8


def test_a_longer_call_after_a_shorter_one_and_back(tiny_pair):
    # the workspace is grow-only and reused: rows past a level's valid length must read as zeros again
    _, _, oracle, enc = tiny_pair
    for samples in (500, 3000, 333, 3000):
        a = synth_audio(2, samples, samples)
        margins = []
        want = oracle.encode(a, margins=margins).numpy()
        assert count_near_tie_frames(want, enc.encode(a), [m.numpy() for m in margins], tol=1e-4) <= 2


def test_encoder_errors(tiny_pair):
    d, _, _, enc = tiny_pair
    with pytest.raises(q.AudioDecodingFailed):
        enc.encode(np.zeros((1, 2, 100), np.float32))                  # not mono
    with pytest.raises(q.AudioDecodingFailed):
        enc.encode(np.zeros((1, 0), np.float32))                       # empty audio
    lite = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7), "speech_tokenizer")
    with pytest.raises(q.AudioDecodingFailed) as ei:
        q.Qwen3TTSSpeechTokenizerEncoder(lite)                         # no encoder_config: Qwen3.swift:433
    assert "encoder" in str(ei.value)
    with pytest.raises(q.AudioDecodingFailed):
        q.Qwen3TTSSpeechTokenizerEncoder(d, precision=q.PREC_BF16)     # bf16 pairs carry too few mantissa bits for an argmin


def test_codes_feed_the_decoder(tiny_pair):
    # voice cloning chains encode -> (talker) -> decode on the same tokenizer directory (Qwen3.swift:430-440, 1186)
    d, _, _, enc = tiny_pair
    codes = enc.encode(synth_audio(2, 2400, 5))
    tok = q.Qwen3TTSSpeechTokenizer(d, precision=q.PREC_FP32)
    pcm = tok.decoder(codes)
    assert pcm.shape == (2, 1, codes.shape[2] * tok.config.total_upsample) and np.isfinite(pcm).all()
    tok.close()
