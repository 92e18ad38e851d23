"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes mirror of the reference's
Swift API), against the CPU oracle on the same seeded inputs and against the committed golden
fixtures.  Tolerances are north_star's: bit-exact codebook indexing / dequantised sums and PCM
max-abs <= 1e-4 in fp32 mode; SNR >= 40 dB in the 16-bit tensor-core mode."""
import os

import numpy as np
import pytest
import torch

import qwen3tts_cuda as q
from oracle import decoder as od
from oracle import weights as ow
from tools.fixtures import checkpoint_dir
from tools.q3cfg import GOLDEN_CODES_5x16
from tools.synth_checkpoint import synth_codes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def tiny_tok(tiny_dir):
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32)
    yield tok
    tok.close()


@pytest.fixture(scope="module")
def full_tok(full_dir):
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP32)
    yield tok
    tok.close()


def _nct_codes(cfg, B, T, seed, **kw):
    return synth_codes(cfg, B, T, seed, **kw)


def test_native_library_is_loaded_and_sees_the_gpu():
    assert q.device_count() >= 1
    assert os.path.exists(q.library_path())


# ---- fp32 mode: tiny architecture, every stage -----------------------------------------------------
@pytest.mark.parametrize("B,T", [(1, 1), (1, 6), (2, 9), (3, 33)])
def test_fp32_stage_taps_match_oracle_tiny(tiny_tok, tiny_oracle, B, T):
    cfg, w, dec = tiny_oracle
    codes = _nct_codes(cfg, B, T, 100 + T)
    taps = {}
    ref = dec.forward(codes, taps).numpy()
    tiny_tok.set_taps(True)
    out = tiny_tok.decoder(codes)
    assert out.shape == ref.shape
    # bit-exact dequantised sums (sequential fp32 adds, ST.swift:84-93) -- compare against an fp32 oracle
    taps32 = {}
    od.OracleDecoder(cfg, w, torch.float32).forward(codes, taps32)
    for name in ("rvq_sum_first", "rvq_sum_rest"):
        got = tiny_tok.stage_tap(name)
        assert np.array_equal(got, taps32[name].numpy()), name
    for name in ("quantized", "pre_conv", "pre_transformer", "upsample0", "upsample1", "init_conv",
                 "block0", "block1", "block2", "block3", "out_conv"):
        got = tiny_tok.stage_tap(name)
        want = taps[name].numpy()
        assert got.shape == want.shape, name
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= 2e-5 * scale, (name, float(np.abs(got - want).max()))
    tiny_tok.set_taps(False)
    assert np.abs(out - ref).max() <= FP32_TOL


def test_fp32_tiny_golden_fixture(tiny_tok):
    gold = np.load(os.path.join(GOLD, "golden_tiny.npz"))
    out = tiny_tok.decoder(gold["codes"])
    assert np.abs(out - gold["reference_audio"]).max() <= FP32_TOL


def test_fp32_causal_sw_mode_matches_oracle(tiny_dir, tiny_oracle):
    cfg, w, _ = tiny_oracle
    tok = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, attn_mode=q.ATTN_CAUSAL_SW)
    codes = _nct_codes(cfg, 2, 11, 77)
    ref = od.OracleDecoder(cfg, w, torch.float64, attn_mode="causal_sw").forward(codes).numpy()
    out = tok.decoder(codes)
    assert np.abs(out - ref).max() <= FP32_TOL
    gold = np.load(os.path.join(GOLD, "golden_tiny.npz"))
    assert np.abs(tok.decoder(gold["codes"]) - gold["causal_sw_audio"]).max() <= FP32_TOL
    tok.close()


# ---- the reference's API surface ---------------------------------------------------------------------
def test_decode_api_lengths_and_layouts(tiny_tok, tiny_oracle):
    cfg, w, dec = tiny_oracle
    codes = _nct_codes(cfg, 3, 8, 5, zero_frac=0.3)       # zeros sprinkled in codebook 0
    bt16 = np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))
    ref_audio, ref_len = dec.decode(bt16)
    audio, lengths = tiny_tok.decode(bt16)
    assert audio.shape == (3, 8 * cfg.total_upsample) and audio.dtype == np.float32
    assert np.array_equal(lengths, ref_len) and lengths.dtype == np.int32
    assert np.abs(audio - ref_audio.numpy()).max() <= FP32_TOL
    nct = tiny_tok.decoder(codes)                           # [B,16,T] -> [B,1,S]
    assert nct.shape == (3, 1, 8 * cfg.total_upsample)
    assert np.array_equal(nct[:, 0, :], audio)              # both entry points run the same chain


def test_batch_rows_equal_single_decodes(tiny_tok, tiny_oracle):
    cfg, _, _ = tiny_oracle
    codes = _nct_codes(cfg, 4, 7, 8)
    full = tiny_tok.decoder(codes)
    for b in range(4):
        one = tiny_tok.decoder(codes[b:b + 1])
        assert np.array_equal(full[b], one[0])              # same kernels, same order: bit-identical


def test_varlen_equals_per_utterance_decode(tiny_tok, tiny_oracle):
    cfg, w, dec = tiny_oracle
    rng = np.random.default_rng(3)
    lens = [1, 17, 5, 0, 9, 17, 2]
    utts = [np.ascontiguousarray(np.transpose(_nct_codes(cfg, 1, max(L, 1), 200 + i)[0], (1, 0)))[:L] for i, L in enumerate(lens)]
    pcms, lengths = tiny_tok.decode_varlen(utts)
    assert [p.shape[0] for p in pcms] == [L * cfg.total_upsample for L in lens]
    for i, (u, L) in enumerate(zip(utts, lens)):
        if L == 0:
            assert lengths[i] == 0
            continue
        ref_audio, ref_len = dec.decode(u[None])
        assert np.abs(pcms[i] - ref_audio.numpy()[0]).max() <= FP32_TOL, i
        assert lengths[i] == ref_len[0]
        single, _ = tiny_tok.decode(u[None])
        assert np.abs(pcms[i] - single[0]).max() <= 1e-6      # padding never leaks into valid frames (H5)


def test_int16_decode_equals_host_conversion_of_float_decode(tiny_tok, tiny_oracle):
    # row N1 of SURVEY 8(f): Int16(clamp(x,-1,1) * 32767) (main.swift:158-160) written by the tail kernel itself;
    # bit-exact against the C ABI's host conversion and against numpy's truncation of the oracle-checked float PCM
    cfg, _, _ = tiny_oracle
    codes = np.ascontiguousarray(np.transpose(_nct_codes(cfg, 3, 9, 77), (0, 2, 1)))        # [B,T,16]
    audio, lengths = tiny_tok.decode(codes)
    audio16, lengths16 = tiny_tok.decode_int16(codes)
    assert audio16.dtype == np.int16 and audio16.shape == audio.shape
    assert np.array_equal(lengths, lengths16)
    assert np.array_equal(audio16.ravel(), q.pcm_to_int16(audio.ravel()))
    assert np.array_equal(audio16, np.trunc(np.clip(audio, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16))
    utts = [codes[0, :4], codes[1, :0], codes[2]]
    pcms, _ = tiny_tok.decode_varlen(utts)
    pcms16, _ = tiny_tok.decode_varlen(utts, int16=True)
    for f, i in zip(pcms, pcms16):
        assert np.array_equal(i, q.pcm_to_int16(f))
    again, _ = tiny_tok.decode(codes)                       # the float path is untouched by the int16 call before it
    assert np.array_equal(again, audio)


def test_int16_decode_16bit_engine(full_dir, full_oracle):
    cfg, _, _ = full_oracle
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16)
    codes = np.ascontiguousarray(np.transpose(_nct_codes(cfg, 2, 11, 5), (0, 2, 1)))
    audio, _ = tok.decode(codes)
    audio16, _ = tok.decode_int16(codes)
    assert np.array_equal(audio16.ravel(), q.pcm_to_int16(audio.ravel()))
    tok.close()


def test_cuda_graph_replay_is_bit_identical(full_dir, full_oracle):
    # small decodes are launch-bound: the launch chain is captured on its second sighting and replayed afterwards.
    # Same kernels in the same order => bit-identical PCM, also when the codes (and so the lengths) change under the same shape.
    cfg, _, _ = full_oracle
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16)
    c1 = np.ascontiguousarray(np.transpose(_nct_codes(cfg, 2, 13, 71), (0, 2, 1)))
    c2 = np.ascontiguousarray(np.transpose(_nct_codes(cfg, 2, 13, 72, zero_frac=0.3), (0, 2, 1)))
    tok.set_graphs(0)
    ref1, len1 = tok.decode(c1)
    ref2, len2 = tok.decode(c2)
    n0 = tok.launch_count()
    tok.decode(c1)
    per_call = tok.launch_count() - n0
    tok.set_graphs(1)
    for i, (c, ref, ln) in enumerate([(c1, ref1, len1), (c1, ref1, len1), (c2, ref2, len2), (c1, ref1, len1), (c2, ref2, len2)]):
        n0 = tok.launch_count()
        out, lengths = tok.decode(c)          # eager, capture + launch, replay, replay, replay
        assert np.array_equal(out, ref) and np.array_equal(lengths, ln), i
        assert tok.launch_count() - n0 == per_call, i
    i16, _ = tok.decode_int16(c1)             # another output format = another graph key
    i16b, _ = tok.decode_int16(c1)
    assert np.array_equal(i16, i16b) and np.array_equal(i16.ravel(), q.pcm_to_int16(ref1.ravel()))
    tok.set_graphs(-1)
    out, _ = tok.decode(c2)                   # automatic mode: 26 frames <= 2048
    assert np.array_equal(out, ref2)
    tok.close()


def test_microbatching_is_invisible(tiny_dir, tiny_oracle):
    cfg, _, _ = tiny_oracle
    codes = _nct_codes(cfg, 5, 12, 31)
    a = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32)
    b = q.Qwen3TTSSpeechTokenizer(tiny_dir, precision=q.PREC_FP32, max_frames_per_launch=25)   # 2 utterances per launch
    assert np.array_equal(a.decoder(codes), b.decoder(codes))
    a.close()
    b.close()


def test_empty_and_error_inputs(tiny_tok, tiny_oracle):
    cfg, _, _ = tiny_oracle
    out = tiny_tok.decoder(np.zeros((0, cfg.num_quantizers, 4), np.int32))
    assert out.shape == (0, 1, 4 * cfg.total_upsample)
    out = tiny_tok.decoder(np.zeros((2, cfg.num_quantizers, 0), np.int32))
    assert out.shape == (2, 1, 0)
    pcms, lengths = tiny_tok.decode_varlen([])
    assert pcms == [] and lengths.shape == (0,)
    bad = _nct_codes(cfg, 1, 4, 1)
    bad[0, 3, 2] = cfg.codebook_size                        # acoustic id out of range
    with pytest.raises(q.AudioDecodingFailed) as e:
        tiny_tok.decoder(bad)
    assert e.value.status == 1
    ok = _nct_codes(cfg, 1, 4, 1)
    ok[0, 0, 1] = cfg.semantic_codebook_size - 1            # semantic ids may exceed codebook_size (4096 table)
    tiny_tok.decoder(ok)
    with pytest.raises(q.AudioDecodingFailed):
        tiny_tok.decoder(np.zeros((1, 3, 4), np.int32))     # wrong number of codebooks
    with pytest.raises(q.AudioDecodingFailed):
        tiny_tok.stage_tap("no_such_stage")


@pytest.mark.parametrize("kw", [dict(dtype="float16"), dict(with_encoder_stub=True), dict(mlx_layout=True)])
def test_checkpoint_variants_decode_identically(tiny_cfg, tiny_oracle, kw):
    cfg, w, dec = tiny_oracle
    st = os.path.join(checkpoint_dir(tiny_cfg, seed=7, **kw), "speech_tokenizer")
    tok = q.Qwen3TTSSpeechTokenizer(st, precision=q.PREC_FP32)
    codes = _nct_codes(cfg, 1, 6, 12)
    out = tok.decoder(codes)
    if kw.get("dtype") == "float16":     # the lite variant: fp16 on disk: compare against an oracle on the same file
        c2, w2 = ow.load_decoder(st)
        ref = od.OracleDecoder(c2.decoder_config, w2, torch.float64).forward(codes).numpy()
    else:
        ref = dec.forward(codes).numpy()
    assert np.abs(out - ref).max() <= FP32_TOL
    tok.close()


def test_device_pointer_path(tiny_tok, tiny_oracle):
    cfg, _, _ = tiny_oracle
    codes = _nct_codes(cfg, 2, 10, 19)
    want = tiny_tok.decoder(codes)
    d_codes = torch.from_numpy(codes).cuda()
    d_pcm = torch.empty((2, 10 * cfg.total_upsample), dtype=torch.float32, device="cuda")
    d_len = torch.empty(2, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    tiny_tok.decode_device(d_codes.data_ptr(), 2, 10, d_pcm.data_ptr(), d_len.data_ptr(), s.cuda_stream)
    tiny_tok.sync(s.cuda_stream)
    assert np.array_equal(d_pcm.cpu().numpy(), want[:, 0, :])
    assert d_len.cpu().tolist() == [10 * cfg.total_upsample] * 2


# ---- full-size architecture ----------------------------------------------------------------------------
def test_full_golden_grid_fp32(full_tok, full_oracle):
    cfg, w, dec = full_oracle
    gold = np.load(os.path.join(GOLD, "golden_full_5x16.npz"))
    codes = np.asarray(GOLDEN_CODES_5x16, dtype=np.int32)[None]          # [1,5,16]  Tests.swift:37-47
    full_tok.set_taps(True)
    audio, lengths = full_tok.decode(codes)
    assert audio.shape == (1, 9600) and lengths.tolist() == [9600]       # Tests.swift:253; ST.swift:833
    assert np.abs(audio - gold["audio"]).max() <= FP32_TOL
    # the reference test's stage walk (Tests.swift:57-257): shapes and statistics per stage
    for name, shp in (("quantized", (1, 512, 5)), ("pre_conv", (1, 1024, 5)), ("pre_transformer", (1, 1024, 5)),
                      ("upsample0", (1, 1024, 10)), ("upsample1", (1, 1024, 20)), ("init_conv", (1, 1536, 20)),
                      ("block0", (1, 768, 160)), ("block1", (1, 384, 800)), ("block2", (1, 192, 3200)),
                      ("block3", (1, 96, 9600)), ("out_conv", (1, 1, 9600))):
        t = full_tok.stage_tap(name)
        assert t.shape == shp, name
        assert abs(float(t.std()) - gold[f"stat_{name}"][0]) <= 1e-3 * max(1.0, float(t.std())), name
    assert np.abs(full_tok.stage_tap("quantized")[0, :10, 0] - gold["quantized_0_10_0"]).max() <= 1e-4
    assert full_tok.weight_shape("decoder.decoder.initConv.conv.weight") == (1536, 7, 1024)   # Tests.swift:131-132
    full_tok.set_taps(False)


def test_full_fp32_vs_oracle_and_rvq_bit_exact(full_tok, full_oracle):
    cfg, w, dec = full_oracle
    codes = _nct_codes(cfg, 2, 20, 1001)
    taps = {}
    ref = dec.forward(codes, taps).numpy()
    full_tok.set_taps(True)
    out = full_tok.decoder(codes)
    assert np.array_equal(full_tok.stage_tap("rvq_sum_first"), taps["rvq_sum_first"].numpy())
    assert np.array_equal(full_tok.stage_tap("rvq_sum_rest"), taps["rvq_sum_rest"].numpy())
    full_tok.set_taps(False)
    assert np.abs(out - ref).max() <= FP32_TOL
    assert float(np.abs(ref).max()) < 1.0


@pytest.mark.parametrize("prec,floor", [(q.PREC_FP16, 40.0), (q.PREC_BF16, 22.0)])
def test_full_16bit_snr(full_dir, full_oracle, prec, floor):
    # north_star: SNR >= 40 dB in the 16-bit mode.  fp16 operands meet it; bf16 operands cannot on this
    # decoder (27 dB even with ideal fp32 accumulation -- see DESIGN.md "precision"), so bf16 is
    # checked against its own documented floor.
    cfg, w, dec = full_oracle
    codes = _nct_codes(cfg, 2, 20, 1001)
    ref = dec.forward(codes).numpy()
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=prec)
    out = tok.decoder(codes)
    snr = od.snr_db(ref, out)
    print(f"precision {prec}: SNR {snr:.1f} dB")
    assert snr >= floor
    tok.close()


def test_full_size_properties_config1(full_dir):
    # BASELINE config 1 (B=1, T=125) and a slice of config 2: size-independent properties --
    # batch rows == single decodes, prefix-invariance of everything but attention is covered at tiny
    # size; here: determinism, range, lengths.
    from tools.q3cfg import DecoderConfig
    cfg = DecoderConfig()
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16)
    codes = _nct_codes(cfg, 2, 125, 1001)
    bt16 = np.ascontiguousarray(np.transpose(codes, (0, 2, 1)))
    a1, l1 = tok.decode(bt16)
    a2, l2 = tok.decode(bt16)
    assert a1.shape == (2, 240000) and np.array_equal(a1, a2) and l1.tolist() == [240000, 240000]
    assert np.isfinite(a1).all() and np.abs(a1).max() <= 1.0 and a1.std() > 0.01
    single, _ = tok.decode(bt16[1:2])
    assert np.abs(single[0] - a1[1]).max() <= 1e-6
    tok.close()


def test_device_decodes_on_two_streams_share_the_workspace_safely(full_dir, full_oracle):
    # q3tts_decode_device is asynchronous on the CALLER's stream while arena / metadata / error flag belong to the model: two
    # back-to-back decodes on different streams must be ordered on the device (chain_done event), not just on the host.
    cfg, _, _ = full_oracle
    tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16)
    ca, cb = _nct_codes(cfg, 4, 60, 11), _nct_codes(cfg, 3, 45, 12)      # different shapes: metadata and plan change between the calls
    want_a, want_b = tok.decoder(ca)[:, 0], tok.decoder(cb)[:, 0]
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    da, db = torch.from_numpy(ca).cuda(), torch.from_numpy(cb).cuda()
    pa = torch.empty((4, 60 * cfg.total_upsample), dtype=torch.float32, device="cuda")
    pb = torch.empty((3, 45 * cfg.total_upsample), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        pa.zero_(); pb.zero_()
        torch.cuda.synchronize()
        tok.decode_device(da.data_ptr(), 4, 60, pa.data_ptr(), 0, sa.cuda_stream)
        tok.decode_device(db.data_ptr(), 3, 45, pb.data_ptr(), 0, sb.cuda_stream)   # enqueued while A is still running
        tok.decode_device(da.data_ptr(), 4, 60, pa.data_ptr(), 0, 0)                 # and once more on the legacy default stream
        tok.sync(sb.cuda_stream)
        tok.sync(0)
        torch.cuda.synchronize()
        assert np.array_equal(pa.cpu().numpy(), want_a)
        assert np.array_equal(pb.cpu().numpy(), want_b)
    tok.close()


def test_two_models_on_one_device_do_not_share_kernel_state(tiny_cfg):
    # the tiny architecture's last block has 72 channels: the 16-bit tail fallback, whose weights used to live in ONE process-wide
    # __constant__ symbol -- two models interleaved on different streams must each decode with their own outConv
    dirs = [os.path.join(checkpoint_dir(tiny_cfg, seed=sd), "speech_tokenizer") for sd in (7, 8)]
    toks = [q.Qwen3TTSSpeechTokenizer(d, precision=q.PREC_FP16) for d in dirs]
    codes = _nct_codes(tiny_cfg, 6, 40, 3)
    want = [t.decoder(codes) for t in toks]
    assert not np.array_equal(want[0], want[1])
    d_codes = torch.from_numpy(codes).cuda()
    outs = [torch.empty((6, 40 * tiny_cfg.total_upsample), dtype=torch.float32, device="cuda") for _ in toks]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for _ in range(5):
        for t, o, s in zip(toks, outs, streams):
            t.decode_device(d_codes.data_ptr(), 6, 40, o.data_ptr(), 0, s.cuda_stream)
    torch.cuda.synchronize()
    for t, o, w in zip(toks, outs, want):
        assert np.array_equal(o.cpu().numpy(), w[:, 0])
        t.close()


@pytest.mark.skipif(q.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices(full_dir, full_oracle):
    # one handle per GPU.  The >48 KB dynamic-smem opt-in of the tcgen05 kernels is per device.
    cfg, _, _ = full_oracle
    codes = _nct_codes(cfg, 2, 30, 21)
    outs = []
    for dev in (0, 1):
        tok = q.Qwen3TTSSpeechTokenizer(full_dir, precision=q.PREC_FP16, device=dev)
        outs.append(tok.decoder(codes))
        tok.close()
    assert np.array_equal(outs[0], outs[1])
