"""Kernel-level GPU tests through the C ABI's test hooks: the tcgen05 multi-tap GEMM (every epilogue mode the
decoder blocks use, ragged utterance lengths, partial K blocks, resident and streamed weights) and the fused
residual-unit kernel, each against the CUDA-core kernel(s) of the same op on the same seeded random data.
Tolerance: the two paths round the same fp32 values to 16 bits after different summation orders, so the outputs
may differ by one or two units in the last place of a half at |x| <= 8 (2^-8 * 2)."""
import pytest

import qwen3tts_cuda as q

pytestmark = pytest.mark.gpu
ULP = 2.0 ** -8          # fp16 ulp at magnitude [4, 8)


@pytest.mark.parametrize("name,cin,n,taps,dil,mode", [
    ("b3.conv7.d1", 96, 96, 7, 1, 0), ("b3.conv7.d9", 96, 96, 7, 9, 0), ("b3.conv1", 96, 96, 1, 1, 1),
    ("b3.convT", 192, 288, 2, 1, 2), ("b2.conv7.d3", 192, 192, 7, 3, 0), ("b2.conv1", 192, 192, 1, 1, 1),
    ("b1.conv7.d9", 384, 384, 7, 9, 0), ("b1.convT", 768, 1920, 2, 1, 2), ("b0.conv1", 768, 768, 1, 1, 1),
    ("plain", 512, 1024, 3, 1, 3),
])
def test_tcgen05_gemm_matches_cuda_core_gemm(name, cin, n, taps, dil, mode):
    # 3 utterances of 1500 / 1463 / 1426 rows: partial last tiles, tiles past an utterance's length, pair tiles
    _, dy, da = q.debug_conv_gemm(3, 1500, cin, n, taps, dil, mode, q.PREC_FP16, 0)
    assert dy <= 2 * ULP and da <= 2 * ULP, (name, dy, da)


def test_tcgen05_gemm_resident_weights_many_tiles_per_cta():
    # enough rows that every CTA walks several tiles with the weights resident in smem
    _, dy, da = q.debug_conv_gemm(2, 200_000, 96, 96, 7, 3, 0, q.PREC_FP16, 0)
    assert da <= 2 * ULP
    _, dy, da = q.debug_conv_gemm(2, 200_000, 96, 96, 1, 1, 1, q.PREC_FP16, 0)
    assert dy <= 2 * ULP and da <= 2 * ULP


@pytest.mark.parametrize("dil", [1, 3, 9])
@pytest.mark.parametrize("out_snake", [0, 1])
def test_fused_residual_unit_matches_composed_unit(dil, out_snake):
    # strips that start mid-utterance, utterance boundaries inside a strip, dummy tiles at the end of the last strips
    for B, rows in ((3, 20_000), (1, 100), (5, 1000)):
        _, d = q.debug_resunit(B, rows, dil, out_snake, q.PREC_FP16, 0)
        assert d <= 3 * ULP, (B, rows, dil, out_snake, d)


def test_fused_residual_unit_bf16():
    _, d = q.debug_resunit(2, 5000, 3, 0, q.PREC_BF16, 0)
    assert d <= 0.13          # bf16: 8 mantissa bits, |x| <= 8


@pytest.mark.parametrize("C", [192, 128, 64])
@pytest.mark.parametrize("dil", [1, 9])
@pytest.mark.parametrize("mode", [0, 1, 2])     # stream only / stream + next operand / operand only (a block's last unit)
def test_gemm_fused_conv7_conv1_unit_matches_composed_unit(C, dil, mode):
    # conv7 -> SnakeBeta -> (operand in tensor memory) -> conv1 -> + residual in ONE tcgen05 kernel vs the two CUDA-core GEMMs;
    # ragged lengths (rows - 211 b), partial last tiles, pair tiles that straddle utterances
    for B, rows in ((3, 3000), (1, 77), (4, 640)):
        _, dy, da = q.debug_fused_unit(B, rows, C, dil, mode, q.PREC_FP16, 0)
        assert dy <= (0 if mode == 2 else 2 * ULP) and da <= 3 * ULP, (C, dil, mode, B, rows, dy, da)


def test_gemm_fused_unit_bf16():
    _, dy, da = q.debug_fused_unit(2, 5000, 192, 3, 1, q.PREC_BF16, 0)
    assert dy <= 0.07 and da <= 0.13
