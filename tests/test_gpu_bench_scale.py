"""Parity at the sizes the bench runs (BASELINE.json configs[0] and configs[1]) and for the pieces that only show up at those
sizes: the multi-key-tile online softmax of the 16-bit attention kernel, the fused residual units inside the real launch chain
(stage taps with the fused kernels ON), the 64-utterance slot geometry, and the final clip (ST.swift:781).
The checker is the CPU oracle (oracle/decoder.py, a restatement of SpeechTokenizer.swift); tolerances are north_star's:
PCM max-abs <= 1e-4 in fp32 mode, SNR >= 40 dB in the 16-bit mode."""
import os

import numpy as np
import pytest
import torch

import qwen3tts_cuda as q
from oracle import decoder as od
from oracle import weights as ow
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig
from tools.synth_checkpoint import synth_codes

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
SNR_16BIT = 40.0


@pytest.fixture(scope="module")
def fp16_tok(full_dir):
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16)
    yield tok
    tok.close()


# ---- BASELINE config 1: B = 1, 16 codebooks x 125 frames -> 240 000 samples, seed 1001 -----------------------
def test_config1_full_fp32_and_fp16_vs_oracle(full_dir, full_oracle, fp16_tok):
    cfg, w, _ = full_oracle
    codes = synth_codes(cfg, 1, 125, 1001)
    truth = od.OracleDecoder(cfg, w, torch.float64).forward(codes).numpy()        # float64 restatement = the truth
    assert truth.shape == (1, 1, 240000) and float(np.abs(truth).max()) < 1.0      # no clipping hides errors
    tok32 = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP32)
    out32 = tok32.decoder(codes)
    tok32.close()
    err = float(np.abs(out32 - truth).max())
    out16 = fp16_tok.decoder(codes)
    snr = od.snr_db(truth, out16)
    print(f"config 1: fp32 max-abs {err:.2e}, fp16 SNR {snr:.1f} dB")
    assert err <= FP32_TOL
    assert snr >= SNR_16BIT


# ---- BASELINE config 2: the real [64,16,375] batch (seed 1002), fp16 ---------------------------------------------------------
def test_config2_real_batch_utterances_vs_oracle(full_oracle, fp16_tok):
    cfg, w, dec32 = full_oracle
    codes = synth_codes(cfg, 64, 375, 1002)
    bt16 = np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))
    audio, lengths = fp16_tok.decode(bt16)                                          # one launch chain, 64 slots x 375 frames
    assert audio.shape == (64, 720000) and lengths.tolist() == [720000] * 64
    assert np.isfinite(audio).all()
    for b in (0, 31, 63):
        ref, ref_len = dec32.decode(bt16[b:b + 1])                                  # the utterance's own B = 1 reference decode
        snr = od.snr_db(ref.numpy()[0], audio[b])
        print(f"config 2 utterance {b}: fp16 SNR {snr:.1f} dB")
        assert snr >= SNR_16BIT, (b, snr)
        assert int(ref_len[0]) == int(lengths[b])
    # slot geometry: a row of the batch is bit-identical to its own B = 1 decode on the GPU (same kernels, same tile walk per slot)
    for b in (17, 63):
        single, _ = fp16_tok.decode(bt16[b:b + 1])
        assert np.array_equal(single[0], audio[b]), b


# ---- the attention kernel on its own, beyond one 64-key tile ------------------------------------------------------------------
def _attention_reference(qkv, nh, nkv, hd, lens, row_begin, window, lp):
    """Oracle SDPA (ST.swift:519-525) per utterance on its valid rows, in float64, on the operands rounded like the device's."""
    B, T, _ = qkv.shape
    cfg = DecoderConfig(num_attention_heads=nh, num_key_value_heads=nkv, head_dim=hd, sliding_window=max(window, 1))
    dec = od.OracleDecoder(cfg, {}, torch.float64, attn_mode="causal_sw" if window else "reference")
    x = torch.from_numpy(qkv)
    if lp is not None:
        x = x.to(lp)
    x = x.to(torch.float64)
    out = np.zeros((B, T, nh * hd), dtype=np.float64)
    for b in range(B):
        lo, hi = int(row_begin[b]) if row_begin is not None else 0, int(lens[b]) if lens is not None else T
        if hi <= lo:
            continue
        rows = x[b, lo:hi]
        L = hi - lo
        qh = rows[:, :nh * hd].reshape(1, L, nh, hd).permute(0, 2, 1, 3)
        kh = rows[:, nh * hd:(nh + nkv) * hd].reshape(1, L, nkv, hd).permute(0, 2, 1, 3).repeat_interleave(nh // nkv, 1)
        vh = rows[:, (nh + nkv) * hd:].reshape(1, L, nkv, hd).permute(0, 2, 1, 3).repeat_interleave(nh // nkv, 1)
        o = dec.sdpa(qh, kh, vh).permute(0, 2, 1, 3).reshape(L, nh * hd)
        out[b, lo:hi] = o.numpy()
    return out


@pytest.mark.parametrize("T", [65, 125, 375, 750])
@pytest.mark.parametrize("window", [0, 72])
@pytest.mark.parametrize("nh,nkv", [(16, 16), (16, 4)])
def test_attention_kernel_multi_key_tile(T, window, nh, nkv):
    hd, B = 64, 3
    rng = np.random.default_rng(1000 + T + window + nkv)
    qkv = rng.standard_normal((B, T, (nh + 2 * nkv) * hd)).astype(np.float32)
    qkv[:, :, :(nh + nkv) * hd] *= 1.5                      # score std ~2.25: a peaky softmax whose running max moves between key tiles
    lens = np.array([T, max(1, T - 37), max(1, T // 2 + 1)], dtype=np.int32)     # ragged slots: padded rows must never leak
    for prec, lp, floor in ((q.PREC_FP16, torch.float16, 55.0), (q.PREC_BF16, torch.bfloat16, 38.0), (q.PREC_FP32, None, 100.0)):
        got = q.debug_attention(qkv, nh, nkv, hd, lens=lens, window=window, precision=prec)
        want = _attention_reference(qkv, nh, nkv, hd, lens, None, window, lp)
        for b in range(B):
            snr = od.snr_db(want[b, :lens[b]], got[b, :lens[b]])
            assert snr >= floor, (T, window, nh, nkv, prec, b, snr)
            assert not got[b, lens[b]:].any()                # rows past the utterance are not written


@pytest.mark.parametrize("window", [0, 72])
def test_attention_kernel_row_begin_history_layout(window):
    # the streaming layout: valid rows [row_begin, len) right-aligned behind a KV history (engine.cu run_stream_batch)
    hd, nh, nkv, B, T = 64, 16, 16, 3, 200
    rng = np.random.default_rng(7 + window)
    qkv = rng.standard_normal((B, T, (nh + 2 * nkv) * hd)).astype(np.float32)
    lens = np.array([200, 150, 77], dtype=np.int32)
    beg = np.array([0, 71, 70], dtype=np.int32)
    got = q.debug_attention(qkv, nh, nkv, hd, lens=lens, row_begin=beg, window=window, precision=q.PREC_FP16)
    want = _attention_reference(qkv, nh, nkv, hd, lens, beg, window, torch.float16)
    for b in range(B):
        # query tiles that start before row_begin produce rows nobody reads; the contract covers rows [row_begin, len)
        assert od.snr_db(want[b, beg[b]:lens[b]], got[b, beg[b]:lens[b]]) >= 55.0, b


# ---- stage taps with the fused kernels ON (16-bit mode) ----------------------------------------------------------------------
def test_fp16_stage_taps_on_the_fused_path(full_oracle, fp16_tok):
    cfg, w, dec32 = full_oracle
    codes = synth_codes(cfg, 2, 20, 1001)
    taps = {}
    ref = dec32.forward(codes, taps).numpy()
    plain = fp16_tok.decoder(codes)
    n0 = fp16_tok.launch_count()
    fp16_tok.decoder(codes)
    per_call = fp16_tok.launch_count() - n0
    fp16_tok.set_taps(True)
    try:
        n0 = fp16_tok.launch_count()
        tapped = fp16_tok.decoder(codes)
        tap_launches = fp16_tok.launch_count() - n0
        # The production chain plus 12 tap copies plus the re-run of block 3's last unit without the consumer's activation: were a
        # fused residual unit switched off in tap mode, its block would need three more launches.
        assert tap_launches == per_call + 13, (per_call, tap_launches)
        # same kernels on the same data; only initConv's epilogue differs (it also writes its fp32 tap), so the PCM agrees to rounding
        print("tap-mode PCM bit-identical:", bool(np.array_equal(tapped, plain)))
        assert od.snr_db(plain, tapped) >= 70.0
        worst = {}
        for name in ("quantized", "pre_conv", "pre_transformer", "upsample0", "upsample1", "init_conv",
                     "block0", "block1", "block2", "block3", "out_conv"):
            got = fp16_tok.stage_tap(name)
            want = taps[name].numpy()
            assert got.shape == want.shape, name
            worst[name] = od.snr_db(want, got)
        print("fp16 stage SNR (dB):", {k: round(v, 1) for k, v in worst.items()})
        for name, snr in worst.items():
            assert snr >= SNR_16BIT, (name, snr)
    finally:
        fp16_tok.set_taps(False)
    assert od.snr_db(ref, plain) >= SNR_16BIT


# ---- the clip (ST.swift:781) and the int16 saturation (main.swift:158-160) on samples that really leave [-1, 1] ---------------
def test_clip_is_exercised_fp32_tiny(tiny_cfg):
    st = os.path.join(checkpoint_dir(tiny_cfg, seed=7, out_gain=8.0), "speech_tokenizer")
    c2, w2 = ow.load_decoder(st)
    cfg = c2.decoder_config
    dec = od.OracleDecoder(cfg, w2, torch.float64)
    codes = synth_codes(cfg, 3, 40, 4242)
    taps = {}
    ref = dec.forward(codes, taps).numpy()
    raw = taps["out_conv"].numpy()
    frac = float((np.abs(raw) > 1.0).mean())
    assert frac > 0.05, frac                                   # the fixture drives > 5 % of the samples into the clip
    tok = q.Qwen3TTSSpeechTokenizer(st, precision=q.PREC_FP32)
    out = tok.decoder(codes)
    assert np.abs(out - ref).max() <= FP32_TOL
    assert float(out.max()) == 1.0 and float(out.min()) == -1.0
    sure = np.abs(raw) > 1.0 + 1e-3                            # far enough out that fp32 rounding cannot bring them back
    assert np.array_equal(out[sure], np.sign(raw[sure]).astype(np.float32))
    bt16 = np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))
    a16, _ = tok.decode_int16(bt16)
    af, _ = tok.decode(bt16)
    assert np.array_equal(a16.ravel(), q.pcm_to_int16(af.ravel()))
    assert int(a16.max()) == 32767 and int(a16.min()) == -32767     # Int16(+-1.0 * 32767)
    assert np.array_equal(a16.reshape(raw.shape)[sure], (np.sign(raw[sure]) * 32767).astype(np.int16))
    tok.close()


def test_clip_is_exercised_fp16_full(full_cfg):
    st = os.path.join(checkpoint_dir(full_cfg, out_gain=8.0), "speech_tokenizer")
    c2, w2 = ow.load_decoder(st)
    cfg = c2.decoder_config
    dec = od.OracleDecoder(cfg, w2, torch.float32)
    codes = synth_codes(cfg, 2, 20, 99)
    taps = {}
    ref = dec.forward(codes, taps).numpy()
    raw = taps["out_conv"].numpy()
    assert float((np.abs(raw) > 1.0).mean()) > 0.05
    tok = q.Qwen3TTSSpeechTokenizer(st, precision=q.PREC_FP16)       # the HMMA tail kernel's store_pcm
    out = tok.decoder(codes)
    assert float(out.max()) == 1.0 and float(out.min()) == -1.0 and np.abs(out).max() <= 1.0
    assert od.snr_db(ref, out) >= SNR_16BIT
    sure = np.abs(raw) > 1.05
    assert np.array_equal(out[sure], np.sign(raw[sure]).astype(np.float32))
    bt16 = np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))
    a16, _ = tok.decode_int16(bt16)
    assert np.array_equal(a16.ravel(), q.pcm_to_int16(out.ravel()))
    assert int(a16.max()) == 32767 and int(a16.min()) == -32767
    tok.close()
