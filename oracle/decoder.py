"""torch-CPU restatement of the reference speech-tokenizer decoder (codes -> PCM).

Every function cites the lines of
/root/reference/Sources/Qwen3TTS/Models/SpeechTokenizer.swift (ST.swift) it follows.
Tensors between modules are NCT ``[batch, channels, time]`` exactly like the reference;
MLX ops run NTC internally (ST.swift:294-304) and are restated below with MLX's
documented semantics (conv weight ``[Cout, K, Cin/groups]``, cross-correlation; transposed
conv weight ``[Cout, K, Cin]``, output length ``(L-1)*s + K``).

``attn_mode``: 'reference' = full bidirectional attention, no mask, no RoPE -- what the
reference decoder actually does (ST.swift:763 passes no mask, 512-528 never applies
RoPE).  'causal_sw' = causal attention limited to ``sliding_window`` keys with
optional RoPE: the mode in which chunked streaming is mathematically possible.

``operand`` = 'bf16' rounds the inputs and weights of every conv/linear/attention matmul
to bfloat16 while keeping accumulation, normalisation, activations and the residual
stream in ``dtype``: the numeric model of the CUDA engine's bf16 mode.

TEST INFRASTRUCTURE (see package docstring).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from tools.q3cfg import DecoderConfig

STAGES = ("quantized", "pre_conv", "pre_transformer", "upsample0", "upsample1", "init_conv",
          "block0", "block1", "block2", "block3", "out_snake", "out_conv", "audio")


class OracleDecoder:
    def __init__(self, cfg: DecoderConfig, weights: Dict[str, np.ndarray], dtype=torch.float32,
                 attn_mode: str = "reference", operand: str = "native", use_rope: bool = False,
                 store: str = "native", decode_upsample_rate: int = 0):
        assert attn_mode in ("reference", "causal_sw")
        assert operand in ("native", "bf16", "fp16")
        self.cfg = cfg
        self.dtype = dtype
        self.attn_mode = attn_mode
        self.operand = operand
        self.use_rope = use_rope
        self.decode_upsample_rate = decode_upsample_rate or cfg.total_upsample
        assert store in ("native", "bf16", "fp16")
        self.store = store   # rounding of the residual stream when a stage writes it to HBM
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()
                  if k.startswith("decoder.")}

    # ---- operand rounding (numeric model of the bf16 engine) -------------------------
    def _q(self, t: torch.Tensor) -> torch.Tensor:
        if self.operand == "native":
            return t
        lp = torch.bfloat16 if self.operand == "bf16" else torch.float16
        return t.to(lp).to(self.dtype)

    def _s(self, t: torch.Tensor) -> torch.Tensor:
        if self.store == "native":
            return t
        lp = torch.bfloat16 if self.store == "bf16" else torch.float16
        return t.to(lp).to(self.dtype)

    # ---- MLX op semantics ------------------------------------------------------------
    def conv1d_ntc(self, x, w, b=None, dilation=1, groups=1):
        """MLX conv1d / MLXNN.Conv1d: x [N,T,Cin], w [Cout,K,Cin/groups], no padding."""
        y = F.conv1d(self._q(x).permute(0, 2, 1), self._q(w).permute(0, 2, 1), None, 1, 0, dilation, groups)
        y = y.permute(0, 2, 1)
        return y if b is None else y + b

    def conv_transposed1d_ntc(self, x, w, b, stride):
        """MLXNN.ConvTransposed1d: x [N,L,Cin], w [Cout,K,Cin] -> [N,(L-1)*s+K,Cout]."""
        y = F.conv_transpose1d(self._q(x).permute(0, 2, 1), self._q(w).permute(2, 0, 1), None, stride)
        return y.permute(0, 2, 1) + b

    def linear(self, x, w, b=None):
        y = self._q(x) @ self._q(w).t()
        return y if b is None else y + b

    @staticmethod
    def rms_norm(x, w, eps):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w

    @staticmethod
    def layer_norm(x, w, b, eps):
        mu = x.mean(-1, keepdim=True)
        var = (x - mu).pow(2).mean(-1, keepdim=True)
        return (x - mu) * torch.rsqrt(var + eps) * w + b

    # ---- modules ---------------------------------------------------------------------
    def causal_conv1d(self, x_nct, prefix, k, dilation=1, groups=1):
        """CausalConv1d, ST.swift:293-305: left-pad (k-1)*d zeros, conv, back to NCT."""
        xt = x_nct.permute(0, 2, 1)
        pad = (k - 1) * dilation
        xt = F.pad(xt, (0, 0, pad, 0))
        y = self.conv1d_ntc(xt, self.w[prefix + ".conv.weight"], self.w[prefix + ".conv.bias"], dilation, groups)
        return y.permute(0, 2, 1)

    def causal_transpose_conv1d(self, x_nct, prefix, k, stride):
        """CausalTransposeConv1d, ST.swift:339-353: conv-transpose then drop the last k-s samples."""
        y = self.conv_transposed1d_ntc(x_nct.permute(0, 2, 1), self.w[prefix + ".conv.weight"],
                                       self.w[prefix + ".conv.bias"], stride)
        trim = k - stride
        if trim > 0:
            y = y[:, :-trim, :]
        return y.permute(0, 2, 1)

    def snake_beta(self, x_nct, prefix):
        """SnakeBeta, ST.swift:246-253: x + 1/(e^beta + 1e-9) * sin^2(x * e^alpha)."""
        a = torch.exp(self.w[prefix + ".alpha"]).reshape(1, -1, 1)
        b = torch.exp(self.w[prefix + ".beta"]).reshape(1, -1, 1)
        s = torch.sin(x_nct * a)
        return x_nct + (1.0 / (b + 1e-9)) * (s * s)

    def quantizer_decode(self, codes, taps=None):
        """SplitResidualVectorQuantizer.decode, ST.swift:214-226 (+ 161-169, 81-96, 50-55)."""
        cfg = self.cfg
        ns = cfg.num_semantic_quantizers

        def rvq(part, part_codes):
            q = None
            for i in range(part_codes.shape[1]):          # ST.swift:84-93, sequential left-to-right
                emb = self.w[f"decoder.quantizer.{part}.vq.layers.{i}.codebook.embed.weight"]
                e = emb[part_codes[:, i, :].long()]        # Embedding gather -> [B,T,D]
                e = e.permute(0, 2, 1)                     # ST.swift:54
                q = e if q is None else q + e
            if taps is not None:   # pre-projection sum: the bit-exact part of the dequantiser
                taps["rvq_sum_first" if part == "rvq_first" else "rvq_sum_rest"] = q.detach().clone()
            # Conv1dProjection, ST.swift:110-118 (1x1 conv, no bias)
            w = self.w[f"decoder.quantizer.{part}.output_proj.weight"]
            return self.conv1d_ntc(q.permute(0, 2, 1), w).permute(0, 2, 1)

        quantized = rvq("rvq_first", codes[:, :ns, :])
        if codes.shape[1] > ns:
            quantized = quantized + rvq("rvq_rest", codes[:, ns:, :])
        return quantized

    def _rope(self, q, k):
        # standard rotate-half RoPE (only in 'causal_sw' with use_rope; not in the reference path)
        T, D = q.shape[2], q.shape[3]
        inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, D, 2, dtype=torch.float64) / D))
        ang = torch.arange(T, dtype=torch.float64)[:, None] * inv[None, :]
        cos = torch.cat([ang.cos(), ang.cos()], -1).to(q.dtype)
        sin = torch.cat([ang.sin(), ang.sin()], -1).to(q.dtype)

        def rot(x):
            x1, x2 = x[..., : D // 2], x[..., D // 2:]
            return torch.cat([-x2, x1], -1)
        return q * cos + rot(q) * sin, k * cos + rot(k) * sin

    def attention(self, x, p):
        """DecoderTransformerAttention, ST.swift:512-528."""
        cfg = self.cfg
        B, L, _ = x.shape
        nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        q = self.linear(x, self.w[p + ".q_proj.weight"]).reshape(B, L, nh, hd).permute(0, 2, 1, 3)
        k = self.linear(x, self.w[p + ".k_proj.weight"]).reshape(B, L, nkv, hd).permute(0, 2, 1, 3)
        v = self.linear(x, self.w[p + ".v_proj.weight"]).reshape(B, L, nkv, hd).permute(0, 2, 1, 3)
        if nkv != nh:
            k = k.repeat_interleave(nh // nkv, 1)
            v = v.repeat_interleave(nh // nkv, 1)
        if self.attn_mode == "causal_sw" and self.use_rope:
            q, k = self._rope(q, k)
        o = self.sdpa(q, k, v)
        o = o.permute(0, 2, 1, 3).reshape(B, L, nh * hd)
        return self.linear(o, self.w[p + ".o_proj.weight"])

    def sdpa(self, q, k, v):
        """MLXFast.scaledDotProductAttention(queries:keys:values:scale:mask:) as called at ST.swift:519-525:
        softmax(scale * Q K^T [+ mask]) V, q/k/v [B, heads, L, head_dim], scale = head_dim^-0.5 (ST.swift:502),
        mask nil in the reference (ST.swift:629, 763); 'causal_sw' adds the causal sliding-window mask."""
        L, hd = q.shape[2], q.shape[3]
        scale = float(hd) ** -0.5
        s = (self._q(q) @ self._q(k).transpose(-1, -2)) * scale
        if self.attn_mode == "causal_sw":
            i = torch.arange(L)[:, None]
            j = torch.arange(L)[None, :]
            allowed = (j <= i) & ((i - j) < self.cfg.sliding_window)
            s = s.masked_fill(~allowed, float("-inf"))
        pr = torch.softmax(s, dim=-1)
        return self._q(pr) @ self._q(v)

    def transformer(self, x_ntc):
        """DecoderTransformer, ST.swift:629-643; layer = ST.swift:587-601; MLP = 560-562."""
        cfg = self.cfg
        pt = "decoder.pre_transformer"
        h = self.linear(x_ntc, self.w[pt + ".input_proj.weight"], self.w[pt + ".input_proj.bias"])
        for n in range(cfg.num_hidden_layers):
            p = f"{pt}.layers.{n}"
            r = h
            a = self.rms_norm(h, self.w[p + ".input_layernorm.weight"], cfg.rms_norm_eps)
            a = self.attention(a, p + ".self_attn")
            h = r + a * self.w[p + ".self_attn_layer_scale.scale"]
            r = h
            m = self.rms_norm(h, self.w[p + ".post_attention_layernorm.weight"], cfg.rms_norm_eps)
            g = self.linear(m, self.w[p + ".mlp.gate_proj.weight"])
            u = self.linear(m, self.w[p + ".mlp.up_proj.weight"])
            m = self.linear(F.silu(g) * u, self.w[p + ".mlp.down_proj.weight"])
            h = r + m * self.w[p + ".mlp_layer_scale.scale"]
        h = self.rms_norm(h, self.w[pt + ".norm.weight"], cfg.rms_norm_eps)
        return self.linear(h, self.w[pt + ".output_proj.weight"], self.w[pt + ".output_proj.bias"])

    def convnext(self, x_nct, p):
        """ConvNeXtBlock, ST.swift:385-401 (exact-erf GELU, LayerNorm eps 1e-6)."""
        dim = x_nct.shape[1]
        h = self.causal_conv1d(x_nct, p + ".dwconv", 7, 1, groups=dim)
        h = h.permute(0, 2, 1)
        h = self.layer_norm(h, self.w[p + ".norm.weight"], self.w[p + ".norm.bias"], 1e-6)
        h = self.linear(h, self.w[p + ".pwconv1.weight"], self.w[p + ".pwconv1.bias"])
        h = F.gelu(h)  # exact (erf) form
        h = self.linear(h, self.w[p + ".pwconv2.weight"], self.w[p + ".pwconv2.bias"])
        h = self.w[p + ".gamma"] * h
        return x_nct + h.permute(0, 2, 1)

    def res_unit(self, x, p, dilation):
        """DecoderResidualUnit, ST.swift:430-437."""
        h = self.snake_beta(x, p + ".act1")
        h = self.causal_conv1d(h, p + ".conv1", 7, dilation)
        h = self.snake_beta(h, p + ".act2")
        h = self.causal_conv1d(h, p + ".conv2", 1)
        return self._s(x + h)

    def decoder_block(self, x, p, rate):
        """DecoderBlock, ST.swift:473-480 (dilations 1,3,9 at 468-470)."""
        h = self.snake_beta(x, p + ".snake")
        h = self._s(self.causal_transpose_conv1d(h, p + ".upsample", 2 * rate, rate))
        h = self.res_unit(h, p + ".res1", 1)
        h = self.res_unit(h, p + ".res2", 3)
        h = self.res_unit(h, p + ".res3", 9)
        return h

    # ---- top level -------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, codes_b16t, taps: Optional[dict] = None) -> torch.Tensor:
        """Qwen3TTSSpeechTokenizerDecoder.callAsFunction, ST.swift:754-784.  codes [B,16,T]."""
        cfg = self.cfg
        codes = torch.as_tensor(np.asarray(codes_b16t)).long()

        def tap(name, t):
            if taps is not None:
                taps[name] = t.detach().clone()

        hidden = self.quantizer_decode(codes, taps); tap("quantized", hidden)
        hidden = self.causal_conv1d(hidden, "decoder.pre_conv", 3); tap("pre_conv", hidden)
        hidden = self.transformer(hidden.permute(0, 2, 1)).permute(0, 2, 1); tap("pre_transformer", hidden)
        for i, r in enumerate(cfg.upsampling_ratios):
            hidden = self.causal_transpose_conv1d(hidden, f"decoder.upsample.{i}.0", r, r)
            hidden = self._s(self.convnext(hidden, f"decoder.upsample.{i}.1")); tap(f"upsample{i}", hidden)
        dd = "decoder.decoder"
        wav = self._s(self.causal_conv1d(hidden, dd + ".initConv", 7)); tap("init_conv", wav)
        for i, r in enumerate(cfg.upsample_rates):
            wav = self.decoder_block(wav, f"{dd}.block{i}", r); tap(f"block{i}", wav)
        wav = self.snake_beta(wav, dd + ".outSnake"); tap("out_snake", wav)
        wav = self.causal_conv1d(wav, dd + ".outConv", 7); tap("out_conv", wav)
        wav = torch.clamp(wav, -1.0, 1.0); tap("audio", wav)      # ST.swift:781
        return wav

    @torch.no_grad()
    def decode(self, audio_codes_bt16, taps: Optional[dict] = None):
        """Qwen3TTSSpeechTokenizer.decode, ST.swift:823-836: codes [B,T,16] ->
        (audio [B, 1920*T], audio_lengths [B] = count(code0 > 0) * 1920)."""
        ac = np.asarray(audio_codes_bt16)
        codes = np.transpose(ac, (0, 2, 1))
        wav = self.forward(codes, taps).squeeze(1)
        valid = (ac[:, :, 0] > 0).sum(axis=1)
        lengths = (valid * self.decode_upsample_rate).astype(np.int32)   # ST.swift:833 uses the tokenizer-level rate
        return wav, lengths


def trim_like_generate(audio_1d: np.ndarray, valid_len: int) -> np.ndarray:
    """Caller-side trim, Qwen3.swift:746-752."""
    if 0 < valid_len < audio_1d.shape[0]:
        return audio_1d[:valid_len]
    return audio_1d


def voice_clone_cut(audio_1d: np.ndarray, ref_len: int, total_len: int) -> np.ndarray:
    """Proportional removal of the reference part, Qwen3.swift:1196-1199 (Float arithmetic)."""
    cut = int(np.float32(ref_len) / np.float32(max(total_len, 1)) * np.float32(audio_1d.shape[0]))
    if 0 < cut < audio_1d.shape[0]:
        return audio_1d[cut:]
    return audio_1d


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return 10.0 * math.log10(float((ref ** 2).sum()) / max(float((err ** 2).sum()), 1e-300))
