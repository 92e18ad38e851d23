"""codec-embedding sum for the Talker's next-step input (Qwen3.swift:720-728, 485-491).
CPU: the oracle (sequential adds in the checkpoint dtype) against the float64 definition within the rounding bound.
GPU: the CUDA kernel, through the C ABI, BIT-EXACT against the oracle for bf16 / fp16 / fp32 tables."""
import numpy as np
import pytest
import torch

from oracle import codec_embed as oe
from tools.synth_checkpoint import write_codec_embeddings


def _codes(n, vocab, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.integers(0, v, size=n) for v in vocab], axis=1).astype(np.int32)


@pytest.mark.parametrize("dtype,eps", [("bfloat16", 2.0 ** -8), ("float16", 2.0 ** -11), ("float32", 2.0 ** -24)])
def test_oracle_sequential_sum_is_within_rounding_of_the_definition(tmp_path, dtype, eps):
    d = write_codec_embeddings(str(tmp_path), hidden=64, talker_vocab=48, vocab=32, groups=16, dtype=dtype, seed=5)
    tables = oe.load_tables(d)
    assert len(tables) == 16 and tables[0].shape == (48, 64) and tables[1].shape == (32, 64)
    codes = _codes(200, [48] + [32] * 15, 1)
    out = oe.codec_embed_sum(tables, codes)
    assert out.dtype == tables[0].dtype and out.shape == (200, 64)
    exact = oe.codec_embed_sum_f64(tables, codes)
    # 15 adds, each within half an ulp of a partial sum whose magnitude is bounded by sum |terms|
    bound = 15 * eps * np.abs(np.stack([t.to(torch.float64).numpy()[codes[:, g]] for g, t in enumerate(tables)])).sum(0)
    assert np.all(np.abs(out.to(torch.float64).numpy() - exact) <= bound + 1e-30)
    if dtype != "float32":      # the order matters: a reversed sum rounds differently somewhere
        rev = tables[15][torch.as_tensor(codes[:, 15], dtype=torch.long)]
        for g in range(14, -1, -1):
            rev = rev + tables[g][torch.as_tensor(codes[:, g], dtype=torch.long)]
        assert not torch.equal(rev, out)


def test_single_frame_equals_the_reference_loop(tmp_path):
    # the per-step form (one frame): embedding lookups added one by one, exactly the loop at Qwen3.swift:721-726
    d = write_codec_embeddings(str(tmp_path), hidden=32, talker_vocab=20, vocab=10, groups=16, dtype="bfloat16", seed=9)
    tables = oe.load_tables(d)
    code = _codes(1, [20] + [10] * 15, 2)
    emb = tables[0][int(code[0, 0])]
    for i in range(15):
        emb = emb + tables[1 + i][int(code[0, 1 + i])]
    assert torch.equal(oe.codec_embed_sum(tables, code)[0], emb)


def _to_torch(out, prec):
    import qwen3tts_cuda as q
    if prec == q.PREC_BF16:
        return torch.from_numpy(out.astype(np.int32) << 16).view(torch.float32).to(torch.bfloat16)   # exact: the low bits are zero
    return torch.from_numpy(out)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float32"])
def test_cuda_codec_embed_sum_is_bit_exact(tmp_path, dtype):
    import qwen3tts_cuda as q
    hidden, v0, v, G = 256, 96, 64, 16
    d = write_codec_embeddings(str(tmp_path), hidden=hidden, talker_vocab=v0, vocab=v, groups=G, dtype=dtype, seed=11)
    tables = oe.load_tables(d)
    emb = q.CodecEmbedder(d)
    assert (emb.hidden, emb.groups, emb.vocab) == (hidden, G, [v0] + [v] * (G - 1))
    for n, seed in ((1, 1), (7, 2), (1000, 3)):      # one AR step, a short prefix, a voice-cloning reference
        codes = _codes(n, emb.vocab, seed)
        got = _to_torch(emb(codes), emb.precision)
        ref = oe.codec_embed_sum(tables, codes)
        assert got.dtype == ref.dtype and torch.equal(got, ref), (dtype, n)
    assert emb(np.zeros((0, G), np.int32)).shape == (0, hidden)
    bad = _codes(3, emb.vocab, 4)
    bad[1, 5] = v                                      # one past the table
    with pytest.raises(q.AudioDecodingFailed):
        emb(bad)
    ok = _codes(3, emb.vocab, 4)                       # the error flag does not stick
    assert torch.equal(_to_torch(emb(ok), emb.precision), oe.codec_embed_sum(tables, ok))
    emb.close()


@pytest.mark.gpu
def test_cuda_codec_embed_full_size_tables(tmp_path):
    import qwen3tts_cuda as q
    d = write_codec_embeddings(str(tmp_path), dtype="bfloat16", seed=3)      # 3072 x 2048 + 15 x 2048 x 2048, bf16 (138 MB)
    tables = oe.load_tables(d)
    emb = q.CodecEmbedder(d)
    codes = _codes(375, emb.vocab, 8)
    assert torch.equal(_to_torch(emb(codes), emb.precision), oe.codec_embed_sum(tables, codes))
    emb.close()
