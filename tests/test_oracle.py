"""CPU tests of the oracle against every checkpoint-independent fixture the reference's tests hold
for this path: the golden 5x16 grid, the stage shape chain, the weight-layout KAT,
the tensor inventory, the layout heuristic -- plus the committed golden outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder, ops_def, weights
from tools.q3cfg import DecoderConfig, GOLDEN_CODES_5x16
from tools.synth_checkpoint import decoder_tensor_specs, synth_codes

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_grid_is_the_references():
    # Tests.swift:37-43: 5 frames x 16 codebooks, all < 2048, column 0 non-zero
    g = np.asarray(GOLDEN_CODES_5x16)
    assert g.shape == (5, 16) and g.max() == 2016 and g.min() == 17 and (g[:, 0] > 0).all()
    assert g[0, 0] == 1342 and g[4, 15] == 1498


def test_tensor_inventory_matches_the_paper():
    # docs/paper.tex:218 (271 decoder tensors), :554 (114.3 M decoder parameters)
    specs = decoder_tensor_specs(DecoderConfig())
    assert len(specs) == 271
    n = sum(int(np.prod(s)) for _, s, _, _ in specs)
    assert abs(n / 1e6 - 114.8) < 0.2


def test_layout_heuristic_table():
    # Qwen3.swift:1246-1260, cases from SURVEY Appendix B
    f = weights.is_mlx_conv_layout
    assert not f((1024, 1, 7)) and f((1024, 7, 1))
    assert not f((1, 96, 7)) and f((1, 7, 96))
    assert not f((96, 96, 1)) and f((96, 1, 96))
    assert not f((1536, 768, 16)) and f((768, 16, 1536))
    assert not f((1536, 1024, 7)) and f((1536, 7, 1024))
    assert f((48, 48, 1))            # the heuristic's blind spot below 65 channels
    assert not f((4, 4))


def test_sanitize_key_map_and_layouts(full_oracle):
    cfg, w, _ = full_oracle
    # Tests.swift:131-132: initConv.conv.weight is (1536, 7, 1024) after load
    assert w["decoder.decoder.initConv.conv.weight"].shape == (1536, 7, 1024)
    assert w["decoder.decoder.block0.upsample.conv.weight"].shape == (768, 16, 1536)
    assert w["decoder.decoder.block3.res3.conv2.conv.weight"].shape == (96, 1, 96)
    assert w["decoder.decoder.outConv.conv.weight"].shape == (1, 7, 96)
    assert w["decoder.upsample.0.0.conv.weight"].shape == (1024, 2, 1024)
    assert w["decoder.upsample.1.1.dwconv.conv.weight"].shape == (1024, 7, 1)
    assert w["decoder.quantizer.rvq_rest.output_proj.weight"].shape == (512, 1, 256)
    assert w["decoder.quantizer.rvq_first.vq.layers.0.codebook.embed.weight"].shape == (4096, 256)
    assert w["decoder.quantizer.rvq_rest.vq.layers.14.codebook.embed.weight"].shape == (2048, 256)
    assert len(w) == 255   # 271 - 32 codebook halves + 16 folded tables


def test_shape_chain_and_golden_output(full_oracle):
    # Tests.swift:69, 84, 119-120, 175, 198, 209, 220, 231, 242, 253; ST.swift:833
    cfg, w, dec = full_oracle
    codes = np.asarray(GOLDEN_CODES_5x16, dtype=np.int32)[None]
    taps = {}
    audio, lengths = dec.decode(codes, taps)
    want = {"quantized": (1, 512, 5), "pre_conv": (1, 1024, 5), "pre_transformer": (1, 1024, 5),
            "upsample0": (1, 1024, 10), "upsample1": (1, 1024, 20), "init_conv": (1, 1536, 20),
            "block0": (1, 768, 160), "block1": (1, 384, 800), "block2": (1, 192, 3200),
            "block3": (1, 96, 9600), "out_snake": (1, 96, 9600), "out_conv": (1, 1, 9600)}
    for k, shp in want.items():
        assert tuple(taps[k].shape) == shp, k
    assert tuple(audio.shape) == (1, 9600) and lengths.tolist() == [9600]
    gold = np.load(os.path.join(GOLD, "golden_full_5x16.npz"))
    assert np.array_equal(gold["codes"], codes)
    assert np.abs(audio.numpy() - gold["audio"]).max() <= 2e-5      # fp32 oracle vs the committed fp64 run
    for k in want:
        std = float(taps[k].std())
        assert abs(std - gold[f"stat_{k}"][0]) <= 1e-3 * max(1.0, std), k
    assert float(np.abs(audio.numpy()).max()) < 1.0                   # no clipping hides errors


def test_oracle_vs_definition_level_ops(tiny_oracle):
    cfg, w, dec = tiny_oracle
    codes = synth_codes(cfg, 2, 7, 11)
    for mode in ("reference", "causal_sw"):
        taps = {}
        decoder.OracleDecoder(cfg, w, torch.float64, attn_mode=mode).forward(codes, taps)
        d = ops_def.forward_def(cfg, w, codes, mode)
        for k in decoder.STAGES:
            assert np.abs(d[k] - taps[k].numpy()).max() < 1e-10, (mode, k)


def test_tiny_golden_fixture(tiny_oracle):
    cfg, w, _ = tiny_oracle
    gold = np.load(os.path.join(GOLD, "golden_tiny.npz"))
    for mode in ("reference", "causal_sw"):
        taps = {}
        decoder.OracleDecoder(cfg, w, torch.float64, attn_mode=mode).forward(gold["codes"], taps)
        for k in decoder.STAGES:
            assert np.abs(taps[k].numpy() - gold[f"{mode}_{k}"]).max() < 1e-5 * max(1.0, float(taps[k].abs().max())), (mode, k)


def test_batch_rows_equal_single_decodes(tiny_oracle):
    cfg, w, dec = tiny_oracle
    codes = synth_codes(cfg, 3, 6, 5)
    full = dec.forward(codes).numpy()
    for b in range(3):
        one = dec.forward(codes[b:b + 1]).numpy()
        assert np.abs(full[b] - one[0]).max() < 1e-12


def test_audio_lengths_count_nonzero_not_prefix(tiny_oracle):
    # ST.swift:831-833: code 0 in codebook 0 is "padding" wherever it occurs
    cfg, w, dec = tiny_oracle
    codes = synth_codes(cfg, 2, 8, 9)
    bt16 = np.transpose(codes, (0, 2, 1)).copy()
    bt16[0, 2, 0] = 0
    bt16[0, 5, 0] = 0
    bt16[1, 7, 0] = 0
    audio, lengths = dec.decode(bt16)
    assert lengths.tolist() == [6 * cfg.total_upsample, 7 * cfg.total_upsample]
    assert audio.shape == (2, 8 * cfg.total_upsample)


def test_caller_side_trims():
    # Q3.swift:746-752 and 1196-1199
    a = np.arange(100, dtype=np.float32)
    assert decoder.trim_like_generate(a, 40).shape[0] == 40
    assert decoder.trim_like_generate(a, 0).shape[0] == 100
    assert decoder.trim_like_generate(a, 100).shape[0] == 100
    assert decoder.voice_clone_cut(a, 1, 4)[0] == 25
    assert decoder.voice_clone_cut(a, 0, 4).shape[0] == 100


def test_causality_of_everything_but_attention(tiny_oracle):
    # appending frames must not change earlier samples in causal_sw mode (the streaming premise)
    cfg, w, _ = tiny_oracle
    dec = decoder.OracleDecoder(cfg, w, torch.float64, attn_mode="causal_sw")
    codes = synth_codes(cfg, 1, 10, 21)
    a = dec.forward(codes).numpy()
    b = dec.forward(codes[:, :, :6]).numpy()
    assert np.abs(a[..., : 6 * cfg.total_upsample] - b).max() < 1e-12
    ref = decoder.OracleDecoder(cfg, w, torch.float64, attn_mode="reference")
    a = ref.forward(codes).numpy()
    b = ref.forward(codes[:, :, :6]).numpy()
    assert np.abs(a[..., : 6 * cfg.total_upsample] - b).max() > 1e-6   # the reference's attention is NOT causal (F1)


@pytest.mark.parametrize("operand,floor", [("fp16", 40.0), ("bf16", 20.0)])
def test_16bit_numeric_model(tiny_oracle, operand, floor):
    cfg, w, dec = tiny_oracle
    codes = synth_codes(cfg, 1, 8, 4)
    ref = dec.forward(codes).numpy()
    lo = decoder.OracleDecoder(cfg, w, torch.float32, operand=operand, store=operand).forward(codes).numpy()
    assert decoder.snr_db(ref, lo) >= floor


def test_full_size_operand_precision_bf16_cannot_meet_40_db(full_oracle):
    # DESIGN.md section 4, as evidence instead of prose: on the FULL-size decoder, rounding only the GEMM / conv / attention OPERANDS to
    # bf16 -- fp32 accumulation, fp32 residual stream, fp32 activations, i.e. better than any bf16 engine can do -- already lands far
    # below north_star's 40 dB; the same experiment with fp16 operands clears it.  The 16-bit headline mode is therefore fp16 (same
    # tensor-core rate, `kind::f16`), and bf16 is reported beside it, not instead of it.
    cfg, w, dec32 = full_oracle
    codes = synth_codes(cfg, 1, 20, 1001)
    ref = dec32.forward(codes).numpy()
    snr = {}
    for operand in ("bf16", "fp16"):
        lo = decoder.OracleDecoder(cfg, w, torch.float32, operand=operand).forward(codes).numpy()
        snr[operand] = decoder.snr_db(ref, lo)
    print(f"full-size operand rounding only: bf16 {snr['bf16']:.1f} dB, fp16 {snr['fp16']:.1f} dB")
    assert snr["bf16"] < 32.0          # measured 27 dB: 13 dB short of the bar with everything else ideal
    assert snr["fp16"] >= 42.0         # measured 44.9 dB
    assert snr["fp16"] - snr["bf16"] >= 12.0   # three mantissa bits = 18 dB in theory
