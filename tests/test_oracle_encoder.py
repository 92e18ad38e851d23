"""CPU tests of the encoder oracle: the restatement against definition-level arithmetic, the reference's
padding rule, its key remap, and the properties the architecture guarantees (frame count, causality).  The reference has no
test or golden vector for ``encode``; parity at the MLX boundary is unpinned (oracle/encoder.py header)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oe
from tools.fixtures import checkpoint_dir
from tools.q3cfg import DecoderConfig, EncoderConfig
from tools.enc_compare import count_near_tie_frames
from tools.synth_checkpoint import encoder_tensor_specs, make_encoder_state, synth_audio


@pytest.fixture(scope="module")
def tiny_enc():
    ec = EncoderConfig.tiny()
    d = os.path.join(checkpoint_dir(DecoderConfig.tiny(), seed=7, encoder_cfg=ec), "speech_tokenizer")
    cfg, w = oe.load_encoder(d)
    return cfg, w, oe.OracleEncoder(cfg, w, torch.float64)


def test_extra_padding_rule_known_answers():
    # STE.swift:115-119 by hand: frames = ceil((L + pad - k) / s + 1); ideal = (frames - 1) s + k - pad
    assert oe.extra_padding(10, 4, 2, 2) == 0          # 10 -> 5 frames exactly
    assert oe.extra_padding(11, 4, 2, 2) == 1          # 11 -> 6 frames, one zero on the right
    assert oe.extra_padding(1, 16, 8, 8) == 7          # a single sample still yields one frame
    assert oe.extra_padding(7, 3, 1, 2) == 0           # stride 1 never pads on the right
    assert oe.extra_padding(1441, 12, 6, 6) == 5


@pytest.mark.parametrize("samples", [1, 47, 48, 49, 1000, 1921])
def test_frame_count_is_the_ceil_chain(tiny_enc, samples):
    cfg, _, enc = tiny_enc
    codes = enc.encode(synth_audio(1, samples, 3))
    assert codes.shape == (1, 16, oe.encode_frames(cfg, samples))
    assert int(codes.min()) >= 0 and int(codes.max()) < cfg.codebook_size


def test_full_config_frame_rate_is_12_5_hz():
    cfg = EncoderConfig()
    assert cfg.seanet_stride == 960 and cfg.downsample_stride == 2
    assert oe.encode_frames(cfg, 24000 * 10) == 125 and oe.encode_frames(cfg, 24000 * 10 + 1) == 126


def test_streamable_conv_matches_the_definition(tiny_enc):
    # strided and dilated cases written out tap by tap (cross-correlation, zeros left, extra zeros right)
    cfg, w, enc = tiny_enc
    rng = np.random.default_rng(0)
    for (cin, cout, k, s, d, L) in [(3, 5, 4, 2, 1, 11), (4, 2, 3, 1, 2, 9), (2, 3, 6, 3, 1, 7)]:
        W = rng.normal(size=(cout, k, cin))
        b = rng.normal(size=(cout,))
        x = rng.normal(size=(1, cin, L))
        enc.w["t.conv.conv.weight"] = torch.from_numpy(W)
        enc.w["t.conv.conv.bias"] = torch.from_numpy(b)
        got = enc.sconv(torch.from_numpy(x), "t", k, stride=s, dil=d).numpy()
        eff = (k - 1) * d + 1
        left = eff - s
        nfr = -(-L // s)
        want = np.zeros((1, cout, nfr))
        for t in range(nfr):
            for j in range(k):
                src = t * s + j * d - left
                if 0 <= src < L:
                    want[0, :, t] += W[:, j, :] @ x[0, :, src]
            want[0, :, t] += b
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12


def test_sanitize_rules_produce_the_swift_module_tree():
    cfg = EncoderConfig.tiny()
    raw = {k: v.numpy() for k, v in make_encoder_state(cfg, 5).items()}
    w = oe.sanitize_encoder_weights(raw)
    assert "encoder.encoder.init_conv1d.conv.conv.weight" in w
    assert w["encoder.encoder.init_conv1d.conv.conv.weight"].shape == (cfg.num_filters, cfg.kernel_size, 1)            # forced [o,k,i]
    assert w["encoder.encoder.layers.0.residuals.0.block.0.conv.conv.weight"].shape == (4, 3, 8)
    assert w["encoder.encoder.layers.3.downsample.conv.conv.weight"].shape[1] == 2 * cfg.upsampling_ratios[0]           # ratios reversed
    assert w["encoder.downsample.conv.conv.conv.weight"].shape == (cfg.hidden_size, 2 * cfg.downsample_stride, cfg.hidden_size)
    assert w["encoder.quantizer.rvq_first.input_proj.weight"].shape == (cfg.codebook_dim, 1, cfg.hidden_size)
    assert "encoder.quantizer.rvq_rest.vq.layers.18.codebook.embeddingSum" in w
    assert "encoder.encoder_transformer.transformer.layers.1.gating.linear2.weight" in w
    assert "encoder.encoder_transformer.transformer.layers.0.layer_scale_1.scale" in w
    assert not any("initialized" in k or "semantic_residual" in k or ".mlp." in k for k in w)
    n_raw = len(encoder_tensor_specs(cfg))
    assert len(w) == n_raw - cfg.num_quantizers                        # the `initialized` flags are dropped, nothing else


def test_encode_is_causal_in_whole_frames(tiny_enc):
    cfg, _, enc = tiny_enc
    hop = cfg.seanet_stride * cfg.downsample_stride
    a = synth_audio(1, hop * 12, 11)
    b = a.copy()
    b[..., hop * 7:] = synth_audio(1, hop * 12, 12)[..., hop * 7:]
    ca, cb = enc.encode(a), enc.encode(b)
    assert torch.equal(ca[..., :7], cb[..., :7]) and not torch.equal(ca[..., 7:], cb[..., 7:])


def test_batch_rows_are_independent(tiny_enc):
    _, _, enc = tiny_enc
    a = synth_audio(3, 500, 21)
    all_codes = enc.encode(a)
    for b in range(3):
        assert torch.equal(all_codes[b:b + 1], enc.encode(a[b:b + 1]))


def test_fp32_restatement_agrees_with_fp64_except_at_near_ties(tiny_enc):
    cfg, w, enc64 = tiny_enc
    enc32 = oe.OracleEncoder(cfg, w, torch.float32)
    a = synth_audio(2, 2000, 31)
    margins = []
    c64 = enc64.encode(a, margins=margins).numpy()
    c32 = enc32.encode(a).numpy()
    bad = count_near_tie_frames(c64, c32, [m.numpy() for m in margins], tol=1e-4)
    assert bad <= 0.02 * c64.shape[0] * c64.shape[2]


def test_split_fp16_operands_carry_float32_products():
    # The tensor-core engine's arithmetic, emulated in NumPy: a = hi + lo'/2048 with fp16 hi, lo';
    # [lo' | hi | hi] . [w_hi | w_lo' | 2048 w_hi] / 2048 must reproduce a . w to ~2^-21, and every scaling must be exact.
    rng = np.random.default_rng(5)
    K = 4096
    a = (rng.normal(size=K) * rng.choice([1e-3, 0.05, 1.0, 30.0], size=K)).astype(np.float32)      # activations over several decades
    w = (rng.uniform(-1, 1, size=K) * 0.05).astype(np.float32)                                      # conv weights ~ U(+-1/sqrt(fan_in))

    def split(x):
        hi = x.astype(np.float16)
        lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
        return hi, lo

    a_hi, a_lo = split(a)
    w_hi, w_lo = split(w)
    w_big = (w_hi.astype(np.float32) * np.float32(2048.0)).astype(np.float16)
    assert np.array_equal(w_big.astype(np.float64), w_hi.astype(np.float64) * 2048.0)               # the power-of-two scaling is exact in fp16
    assert np.isfinite(w_big).all() and np.abs(a_lo.astype(np.float32)).min() >= 0
    acc = (a_lo.astype(np.float64) @ w_hi.astype(np.float64) + a_hi.astype(np.float64) @ w_lo.astype(np.float64)
           + a_hi.astype(np.float64) @ w_big.astype(np.float64))
    got = acc / 2048.0
    exact = a.astype(np.float64) @ w.astype(np.float64)
    scale = np.abs(a.astype(np.float64)) @ np.abs(w.astype(np.float64))
    assert abs(got - exact) <= 2.0 ** -20 * scale                                                    # ~22 bits per operand; plain fp16 operands: 2^-10
    plain = a_hi.astype(np.float64) @ w_hi.astype(np.float64)
    assert abs(plain - exact) > 50 * abs(got - exact)
    # per-element reconstruction: hi + lo'/2048 is within 2^-21 relative of the float32 value (lo' stays a NORMAL fp16 number)
    rec = a_hi.astype(np.float64) + a_lo.astype(np.float64) / 2048.0
    nz = np.abs(a) > 1e-4
    assert (np.abs(rec - a)[nz] <= 2.0 ** -21 * np.abs(a)[nz]).all()


def test_transformer_layer_matches_a_definition_level_restatement(tiny_enc):
    # second opinion for the encoder transformer (STE.swift:473-591), written out with explicit loops in float64 and sharing no
    # code with OracleEncoder.transformer: LayerNorm (biased variance), RoPE on pairs (i, i + d/2), causal softmax, layer scale,
    # tanh-GELU MLP
    import math
    cfg, w, enc = tiny_enc
    W = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    rng = np.random.default_rng(3)
    B, T, D = 1, 9, cfg.hidden_size
    x = rng.normal(size=(B, D, T))                                    # NCL, like the reference
    want = enc.transformer(torch.from_numpy(x)).numpy()
    nh = cfg.num_attention_heads
    hd = D // nh
    half = hd // 2

    def layer_norm(v, g, b):
        mu = v.mean()
        var = ((v - mu) ** 2).mean()
        return (v - mu) / math.sqrt(var + 1e-5) * g + b

    def rope(vec, pos):
        out = vec.copy()
        for i in range(half):
            th = pos * cfg.rope_theta ** (-i / half)
            c, s = math.cos(th), math.sin(th)
            out[i] = vec[i] * c - vec[i + half] * s
            out[i + half] = vec[i] * s + vec[i + half] * c
        return out

    h = x[0].T.copy()                                                 # [T, D]
    for li in range(cfg.num_hidden_layers):
        p = f"encoder.encoder_transformer.transformer.layers.{li}"
        n1 = np.stack([layer_norm(h[t], W[f"{p}.norm1.weight"], W[f"{p}.norm1.bias"]) for t in range(T)])
        q = n1 @ W[f"{p}.self_attn.q_proj.weight"].T
        k = n1 @ W[f"{p}.self_attn.k_proj.weight"].T
        v = n1 @ W[f"{p}.self_attn.v_proj.weight"].T
        att = np.zeros((T, D))
        for hh in range(nh):
            sl = slice(hh * hd, (hh + 1) * hd)
            qr = np.stack([rope(q[t, sl], t) for t in range(T)])
            kr = np.stack([rope(k[t, sl], t) for t in range(T)])
            for t in range(T):
                sc = np.array([qr[t] @ kr[u] * hd ** -0.5 for u in range(t + 1)])      # causal: keys 0..t only
                pr = np.exp(sc - sc.max())
                pr /= pr.sum()
                att[t, sl] = pr @ v[: t + 1, sl]
        h = h + W[f"{p}.layer_scale_1.scale"] * (att @ W[f"{p}.self_attn.o_proj.weight"].T)
        n2 = np.stack([layer_norm(h[t], W[f"{p}.norm2.weight"], W[f"{p}.norm2.bias"]) for t in range(T)])
        m = n2 @ W[f"{p}.gating.linear1.weight"].T
        m = m * 0.5 * (1.0 + np.tanh(0.7978845608 * (m + 0.044715 * m ** 3)))
        h = h + W[f"{p}.layer_scale_2.scale"] * (m @ W[f"{p}.gating.linear2.weight"].T)
    assert np.abs(h.T - want[0]).max() < 1e-10


def test_quantizer_matches_a_brute_force_nearest_neighbour_search(tiny_enc):
    # STE.swift:746-759: argmin(|E|^2/2 - x.E^T) is the nearest codebook entry in Euclidean distance; the residual chain by hand
    cfg, w, enc = tiny_enc
    rng = np.random.default_rng(9)
    x = rng.normal(size=(1, cfg.hidden_size, 5)) * 0.7
    got = enc.quantize(torch.from_numpy(x)).numpy()[0]                # [16, 5]
    W = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    row = 0
    for part, n in (("rvq_first", 1), ("rvq_rest", 15)):
        r = x[0].T @ W[f"encoder.quantizer.{part}.input_proj.weight"][:, 0, :].T       # [5, cb]
        for i in range(n):
            b = f"encoder.quantizer.{part}.vq.layers.{i}.codebook"
            E = W[f"{b}.embeddingSum"] / np.maximum(W[f"{b}.clusterUsage"], 1e-5)[:, None]
            for t in range(5):
                d2 = ((E - r[t]) ** 2).sum(-1)
                order = np.argsort(d2)
                if d2[order[1]] - d2[order[0]] > 1e-4:                 # (the oracle searches in float32: skip genuine near-ties)
                    assert got[row, t] == order[0], (part, i, t)
                r[t] = r[t] - E[got[row, t]]
            row += 1
