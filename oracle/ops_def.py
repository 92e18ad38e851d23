"""Definition-level NumPy float64 implementation of the decoder (second opinion).

Shares no code with ``decoder.py`` and calls no library convolution: every op is written
straight from its definition (explicit tap loops + einsum) in MLX's NTC layout.  It exists
to cross-check the fast torch restatement on small shapes.  Same citations as
``decoder.py`` (SpeechTokenizer.swift = ST.swift).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from tools.q3cfg import DecoderConfig

_erf = np.vectorize(math.erf, otypes=[np.float64])


def conv1d(x, w, b=None, dilation=1, groups=1):
    """x [N,T,Cin], w [Cout,K,Cin/groups]; y[n,t,o] = sum_{k,c} x[n,t+k*d,c] w[o,k,c]."""
    N, T, Cin = x.shape
    Cout, K, Cg = w.shape
    To = T - (K - 1) * dilation
    y = np.zeros((N, To, Cout))
    og = Cout // groups
    for g in range(groups):
        xs = x[:, :, g * Cg:(g + 1) * Cg]
        ws = w[g * og:(g + 1) * og]
        for k in range(K):
            y[:, :, g * og:(g + 1) * og] += np.einsum("ntc,oc->nto", xs[:, k * dilation:k * dilation + To, :], ws[:, k, :])
    return y if b is None else y + b


def conv_transposed1d(x, w, b, stride):
    """x [N,L,Cin], w [Cout,K,Cin]; y[n, l*s+k, o] += sum_c x[n,l,c] w[o,k,c]."""
    N, L, Cin = x.shape
    Cout, K, _ = w.shape
    y = np.zeros((N, (L - 1) * stride + K, Cout))
    for l in range(L):
        for k in range(K):
            y[:, l * stride + k, :] += x[:, l, :] @ w[:, k, :].T
    return y + b


def causal_conv(x, w, b, k, d=1, groups=1):      # ST.swift:293-305 (NTC in/out here)
    pad = (k - 1) * d
    xp = np.concatenate([np.zeros((x.shape[0], pad, x.shape[2])), x], axis=1)
    return conv1d(xp, w, b, d, groups)


def causal_tconv(x, w, b, k, s):                 # ST.swift:339-353
    y = conv_transposed1d(x, w, b, s)
    return y[:, : y.shape[1] - (k - s), :] if k > s else y


def snake(x, alpha, beta):                       # ST.swift:246-253 (channels last here)
    s = np.sin(x * np.exp(alpha))
    return x + (1.0 / (np.exp(beta) + 1e-9)) * s * s


def rms_norm(x, w, eps):
    return x / np.sqrt((x * x).mean(-1, keepdims=True) + eps) * w


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def gelu_exact(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def silu(x):
    return x / (1.0 + np.exp(-x))


def forward_def(cfg: DecoderConfig, W: Dict[str, np.ndarray], codes_b16t, attn_mode="reference") -> Dict[str, np.ndarray]:
    """Whole decoder, ST.swift:754-784; returns the stage taps (NCT like the reference)."""
    W = {k: np.asarray(v, dtype=np.float64) for k, v in W.items()}
    codes = np.asarray(codes_b16t).astype(np.int64)
    B, Q, T = codes.shape
    taps = {}
    ns = cfg.num_semantic_quantizers

    def rvq(part, cs):                           # ST.swift:81-96, 161-169
        acc = None
        for i in range(cs.shape[1]):
            e = W[f"decoder.quantizer.{part}.vq.layers.{i}.codebook.embed.weight"][cs[:, i, :]]
            acc = e if acc is None else acc + e
        return conv1d(acc, W[f"decoder.quantizer.{part}.output_proj.weight"])

    h = rvq("rvq_first", codes[:, :ns])
    if Q > ns:
        h = h + rvq("rvq_rest", codes[:, ns:])   # ST.swift:214-226
    taps["quantized"] = h.transpose(0, 2, 1)
    h = causal_conv(h, W["decoder.pre_conv.conv.weight"], W["decoder.pre_conv.conv.bias"], 3)
    taps["pre_conv"] = h.transpose(0, 2, 1)

    pt = "decoder.pre_transformer"               # ST.swift:629-643
    x = h @ W[pt + ".input_proj.weight"].T + W[pt + ".input_proj.bias"]
    nh, hd = cfg.num_attention_heads, cfg.head_dim
    for n in range(cfg.num_hidden_layers):
        p = f"{pt}.layers.{n}"
        a = rms_norm(x, W[p + ".input_layernorm.weight"], cfg.rms_norm_eps)
        q = (a @ W[p + ".self_attn.q_proj.weight"].T).reshape(B, T, nh, hd)
        k = (a @ W[p + ".self_attn.k_proj.weight"].T).reshape(B, T, cfg.num_key_value_heads, hd)
        v = (a @ W[p + ".self_attn.v_proj.weight"].T).reshape(B, T, cfg.num_key_value_heads, hd)
        rep = nh // cfg.num_key_value_heads
        k = np.repeat(k, rep, axis=2)
        v = np.repeat(v, rep, axis=2)
        s = np.einsum("bihd,bjhd->bhij", q, k) * (hd ** -0.5)
        if attn_mode == "causal_sw":
            i = np.arange(T)[:, None]
            j = np.arange(T)[None, :]
            s = np.where((j <= i) & (i - j < cfg.sliding_window), s, -np.inf)
        s = s - s.max(-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(-1, keepdims=True)
        o = np.einsum("bhij,bjhd->bihd", pr, v).reshape(B, T, nh * hd)
        x = x + (o @ W[p + ".self_attn.o_proj.weight"].T) * W[p + ".self_attn_layer_scale.scale"]
        m = rms_norm(x, W[p + ".post_attention_layernorm.weight"], cfg.rms_norm_eps)
        m = (silu(m @ W[p + ".mlp.gate_proj.weight"].T) * (m @ W[p + ".mlp.up_proj.weight"].T)) @ W[p + ".mlp.down_proj.weight"].T
        x = x + m * W[p + ".mlp_layer_scale.scale"]
    x = rms_norm(x, W[pt + ".norm.weight"], cfg.rms_norm_eps)
    h = x @ W[pt + ".output_proj.weight"].T + W[pt + ".output_proj.bias"]
    taps["pre_transformer"] = h.transpose(0, 2, 1)

    for i, r in enumerate(cfg.upsampling_ratios):  # ST.swift:766-775, 385-401
        u = f"decoder.upsample.{i}"
        h = causal_tconv(h, W[u + ".0.conv.weight"], W[u + ".0.conv.bias"], r, r)
        c = causal_conv(h, W[u + ".1.dwconv.conv.weight"], W[u + ".1.dwconv.conv.bias"], 7, 1, groups=h.shape[2])
        c = layer_norm(c, W[u + ".1.norm.weight"], W[u + ".1.norm.bias"], 1e-6)
        c = gelu_exact(c @ W[u + ".1.pwconv1.weight"].T + W[u + ".1.pwconv1.bias"])
        c = c @ W[u + ".1.pwconv2.weight"].T + W[u + ".1.pwconv2.bias"]
        h = h + W[u + ".1.gamma"] * c
        taps[f"upsample{i}"] = h.transpose(0, 2, 1)

    dd = "decoder.decoder"
    h = causal_conv(h, W[dd + ".initConv.conv.weight"], W[dd + ".initConv.conv.bias"], 7)
    taps["init_conv"] = h.transpose(0, 2, 1)
    for i, r in enumerate(cfg.upsample_rates):     # ST.swift:473-480, 430-437
        b = f"{dd}.block{i}"
        h = snake(h, W[b + ".snake.alpha"], W[b + ".snake.beta"])
        h = causal_tconv(h, W[b + ".upsample.conv.weight"], W[b + ".upsample.conv.bias"], 2 * r, r)
        for name, d in (("res1", 1), ("res2", 3), ("res3", 9)):
            p = f"{b}.{name}"
            c = snake(h, W[p + ".act1.alpha"], W[p + ".act1.beta"])
            c = causal_conv(c, W[p + ".conv1.conv.weight"], W[p + ".conv1.conv.bias"], 7, d)
            c = snake(c, W[p + ".act2.alpha"], W[p + ".act2.beta"])
            c = causal_conv(c, W[p + ".conv2.conv.weight"], W[p + ".conv2.conv.bias"], 1)
            h = h + c
        taps[f"block{i}"] = h.transpose(0, 2, 1)
    h = snake(h, W[dd + ".outSnake.alpha"], W[dd + ".outSnake.beta"])
    taps["out_snake"] = h.transpose(0, 2, 1)
    h = causal_conv(h, W[dd + ".outConv.conv.weight"], W[dd + ".outConv.conv.bias"], 7)
    taps["out_conv"] = h.transpose(0, 2, 1)
    taps["audio"] = np.clip(taps["out_conv"], -1.0, 1.0)   # ST.swift:781
    return taps
