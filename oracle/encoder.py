"""CPU restatement of the reference's speech-tokenizer ENCODER (audio -> codes), SURVEY 8(f) row N3.

Follows Sources/Qwen3TTS/Models/SpeechTokenizerEncoder.swift ("STE.swift"):
  115-119   getExtraPaddingForConv1d (Float arithmetic)
  163-186   StreamableConv1d: causal zero padding (left = (k-1)d+1-stride, right = the extra padding)
  262-284   EncoderConv1d (cross-correlation, weight [Cout,K,Cin], + bias)
  333-349   SeanetResnetBlock: x + conv1(elu(conv3(elu(x))))          (true skip)
  384-390   SeanetEncoderLayer: residual blocks, then downsample(elu(.)) with k = 2*ratio, stride = ratio
  436-443   SeanetEncoder: init conv, layers over ratios.reversed(), elu, final conv
  497-526   EncoderAttention: q/k/v/o without bias, RoPE (non-traditional, base = rope_theta) on q and k, SDPA with the mask
  539-541   EncoderMLP: linear2(gelu_tanh(linear1(x))), no bias
  571-590   EncoderTransformerLayer: x + ls1*attn(norm1(x)); x + ls2*mlp(norm2(x)); LayerNorm eps 1e-5
  684-705   EncoderConvDownsample1d: k = 2*stride, no bias (pad mode "edge" is stored but the call pads with zeros)
  731-759   EncoderEuclideanCodebook: E = embed_sum / max(usage, 1e-5); c2 = sum(E^2)/2; argmin(c2 - x E^T) in float32
  816-829   EncoderResidualVectorQuantization.encode: residual -= E[idx] in float32, per layer
  934-941   EncoderSplitResidualVectorQuantizer.encode: rvq_first (1 layer) ++ rvq_rest (nq-1 layers), each with its own input_proj
  1031-1056 Qwen3TTSSpeechTokenizerEncoder.encode: FULL causal mask, first 16 codebooks returned
and the key remap / transposes of Qwen3.swift:1514-1527, 1544-1566, 1590-1679, 1726-1748 (``sanitize_encoder_weights``).

TEST INFRASTRUCTURE (see the package docstring): only tests/ and bench CPU legs import this.  PARITY UNPINNED at the MLX boundary
for the same reason as the decoder (no MLX, no Swift); the reference has no test or golden vector for ``encode``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from tools.q3cfg import EncoderConfig

SEANET_MAPPING = {  # Qwen3.swift:1517-1528
    "encoder.encoder.layers.0.": "encoder.encoder.init_conv1d.",
    "encoder.encoder.layers.1.": "encoder.encoder.layers.0.residuals.0.",
    "encoder.encoder.layers.3.": "encoder.encoder.layers.0.downsample.",
    "encoder.encoder.layers.4.": "encoder.encoder.layers.1.residuals.0.",
    "encoder.encoder.layers.6.": "encoder.encoder.layers.1.downsample.",
    "encoder.encoder.layers.7.": "encoder.encoder.layers.2.residuals.0.",
    "encoder.encoder.layers.9.": "encoder.encoder.layers.2.downsample.",
    "encoder.encoder.layers.10.": "encoder.encoder.layers.3.residuals.0.",
    "encoder.encoder.layers.12.": "encoder.encoder.layers.3.downsample.",
    "encoder.encoder.layers.14.": "encoder.encoder.final_conv1d.",
}


def sanitize_encoder_weights(weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Qwen3.swift:1498-1750 restricted to ``encoder.*`` keys: Swift module paths, conv weights [Cout, K, Cin]."""
    from .weights import is_mlx_conv_layout
    out: Dict[str, np.ndarray] = {}
    codebooks: Dict[str, Dict[str, np.ndarray]] = {}
    for key, value in weights.items():
        if not key.startswith("encoder."):
            continue
        if key.startswith("encoder.quantizer.") and ".codebook." in key:
            base, fld = key.split(".codebook.")
            if fld in ("embed_sum", "cluster_usage"):
                codebooks.setdefault(base, {})[fld] = value
                continue
            if ".initialized" in key:
                continue
        new_key, new_value = key, value
        for py, sw in SEANET_MAPPING.items():   # a dict in the reference: at most one prefix can match a key
            if new_key.startswith(py):
                new_key = new_key.replace(py, sw)
                break
        if ".residuals." in new_key:
            new_key = new_key.replace(".block.1.", ".block.0.").replace(".block.3.", ".block.1.")
        is_seanet_conv = (new_key.startswith("encoder.encoder.") and "encoder_transformer" not in new_key and "quantizer" not in new_key
                          and (".conv.weight" in new_key or ".conv.bias" in new_key))
        if is_seanet_conv:
            new_key = new_key.replace(".conv.weight", ".conv.conv.weight").replace(".conv.bias", ".conv.conv.bias")
            if new_key.endswith(".weight") and value.ndim == 3:
                new_value = np.transpose(value, (0, 2, 1))            # forced: [out, in, k] -> [out, k, in]
        if "encoder.encoder_transformer.layers." in new_key:
            new_key = new_key.replace("encoder.encoder_transformer.layers.", "encoder.encoder_transformer.transformer.layers.")
            new_key = new_key.replace(".input_layernorm.", ".norm1.").replace(".post_attention_layernorm.", ".norm2.")
            new_key = new_key.replace(".mlp.fc1.", ".gating.linear1.").replace(".mlp.fc2.", ".gating.linear2.")
            new_key = new_key.replace(".self_attn_layer_scale.", ".layer_scale_1.").replace(".mlp_layer_scale.", ".layer_scale_2.")
        if new_key.startswith("encoder.downsample.conv.") and "encoder.downsample.conv.conv." not in new_key:
            is_w = new_key.endswith(".weight")
            new_key = new_key.replace("encoder.downsample.conv.", "encoder.downsample.conv.conv.conv.")
            if is_w and value.ndim == 3:
                new_value = np.transpose(value, (0, 2, 1))
        if "encoder.quantizer." in new_key:
            new_key = (new_key.replace(".semantic_residual_vector_quantizer.", ".rvq_first.")
                       .replace(".acoustic_residual_vector_quantizer.", ".rvq_rest."))
            new_key = new_key.replace(".rvq_first.layers.", ".rvq_first.vq.layers.").replace(".rvq_rest.layers.", ".rvq_rest.vq.layers.")
        was_seanet_w = (new_key.startswith("encoder.encoder.") and "encoder_transformer" not in new_key and "quantizer" not in new_key
                        and new_key.endswith(".conv.conv.weight"))
        is_proj = ("input_proj.weight" in new_key or "output_proj.weight" in new_key) and "quantizer" in new_key
        if is_proj and value.ndim == 3:
            new_value = np.transpose(value, (0, 2, 1))
        if "conv.weight" in new_key and value.ndim == 3 and not is_proj and not was_seanet_w:
            if not is_mlx_conv_layout(value.shape):                   # generic branch, from the ORIGINAL value (1696-1700)
                new_value = np.transpose(value, (0, 2, 1))
        out[new_key] = np.ascontiguousarray(new_value)
    for base, data in codebooks.items():                              # 1726-1748: raw sums kept, the module divides (STE.swift:738-743)
        if "cluster_usage" in data and "embed_sum" in data:
            nb = (base.replace(".semantic_residual_vector_quantizer.", ".rvq_first.")
                  .replace(".acoustic_residual_vector_quantizer.", ".rvq_rest."))
            if ".rvq_first.layers." in nb:
                nb = nb.replace(".rvq_first.layers.", ".rvq_first.vq.layers.", 1)
            elif ".rvq_rest.layers." in nb:
                nb = nb.replace(".rvq_rest.layers.", ".rvq_rest.vq.layers.", 1)
            out[f"{nb}.codebook.embeddingSum"] = data["embed_sum"]
            out[f"{nb}.codebook.clusterUsage"] = data["cluster_usage"]
    return out


def load_encoder(speech_tokenizer_dir: str):
    """config.json + weights -> (EncoderConfig, sanitized weight dict).  Raises when the checkpoint has no encoder."""
    import os
    from tools.q3cfg import TokenizerConfig
    from .weights import load_safetensors_dir
    tok = TokenizerConfig.from_json(os.path.join(speech_tokenizer_dir, "config.json"))
    if tok.encoder_config is None:
        raise ValueError("Speech tokenizer encoder not available")     # Qwen3.swift:433
    return EncoderConfig.from_dict(tok.encoder_config), sanitize_encoder_weights(load_safetensors_dir(speech_tokenizer_dir))


def extra_padding(length: int, ksize: int, stride: int, padding_total: int) -> int:
    """STE.swift:115-119, in the reference's Float (fp32) arithmetic."""
    nframes = np.float32(max(length + padding_total - ksize, 0)) / np.float32(stride) + np.float32(1.0)
    ideal = (int(math.ceil(float(nframes))) - 1) * stride + ksize - padding_total
    return max(0, ideal - length)


def encode_frames(cfg: EncoderConfig, samples: int) -> int:
    """Code frames for ``samples`` input samples: every strided conv yields ceil(L / stride) frames."""
    L = samples
    for r in reversed(cfg.upsampling_ratios):
        L = -(-L // r)
    return -(-L // cfg.downsample_stride)


class OracleEncoder:
    def __init__(self, cfg: EncoderConfig, weights: Dict[str, np.ndarray], dtype=torch.float32, valid_quantizers: int = 16):
        self.cfg, self.dtype, self.nvalid = cfg, dtype, valid_quantizers
        self.w = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(dtype) for k, v in weights.items() if k.startswith("encoder.")}
        eps = 1e-5
        self.books = []                                                # (input_proj key, E, c2) in output order
        nrest = min(cfg.num_quantizers - 1, valid_quantizers - 1)
        for part, n in (("rvq_first", 1), ("rvq_rest", nrest)):
            for i in range(n):
                b = f"encoder.quantizer.{part}.vq.layers.{i}.codebook"
                usage = torch.clamp(self.w[f"{b}.clusterUsage"], min=eps)[:, None]
                E = self.w[f"{b}.embeddingSum"] / usage
                self.books.append((part, i, E, (E * E).sum(-1) / 2))

    # ---- building blocks (NCL in / out, like the reference) ----
    def sconv(self, x, prefix, k, stride=1, dil=1, bias=True):
        """StreamableConv1d(causal) -> NormConv1d -> EncoderConv1d"""
        w = self.w[f"{prefix}.conv.conv.weight"].permute(0, 2, 1).contiguous()      # MLX [o,k,i] -> torch [o,i,k]
        b = self.w.get(f"{prefix}.conv.conv.bias") if bias else None
        eff = (k - 1) * dil + 1
        pad_total = eff - stride
        extra = extra_padding(x.shape[-1], eff, stride, pad_total)
        return F.conv1d(F.pad(x, (pad_total, extra)), w, b, stride=stride, dilation=dil)

    @staticmethod
    def elu(x):
        return torch.where(x > 0, x, torch.exp(x) - 1.0)

    def seanet(self, x, taps=None):
        c = self.cfg
        x = self.sconv(x, "encoder.encoder.init_conv1d", c.kernel_size)
        if taps is not None:
            taps["init_conv"] = x
        for li, ratio in enumerate(reversed(c.upsampling_ratios)):
            p = f"encoder.encoder.layers.{li}"
            dil = 1
            for ri in range(c.num_residual_layers):
                r = f"{p}.residuals.{ri}"
                y = self.elu(self.sconv(self.elu(x), f"{r}.block.0", c.residual_kernel_size, dil=dil))
                if taps is not None:
                    taps[f"hid{li}"] = y
                x = self.sconv(y, f"{r}.block.1", 1) + x                             # trueSkip (use_conv_shortcut = false)
                dil *= c.dilation_growth_rate
            if taps is not None:
                taps[f"res{li}"] = x
            x = self.sconv(self.elu(x), f"{p}.downsample", 2 * ratio, stride=ratio)
            if taps is not None:
                taps[f"layer{li}"] = x
        x = self.sconv(self.elu(x), "encoder.encoder.final_conv1d", c.last_kernel_size)
        if taps is not None:
            taps["seanet"] = x
        return x

    def rope(self, x):
        """MLX RoPE(dimensions = head_dim, traditional = false, base): pairs (i, i + d/2).  x [B, H, T, d]"""
        d = x.shape[-1]
        half = d // 2
        freqs = torch.exp(-torch.arange(half, dtype=self.dtype) * (math.log(self.cfg.rope_theta) / half))
        th = torch.arange(x.shape[2], dtype=self.dtype)[:, None] * freqs[None, :]
        cos, sin = torch.cos(th), torch.sin(th)
        x1, x2 = x[..., :half], x[..., half:]
        return torch.cat([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1)

    def transformer(self, x, taps=None):
        c = self.cfg
        h = x.transpose(1, 2)                                                        # NCL -> NLC; no input / output projection (512 == d_model)
        B, T, D = h.shape
        nh, nkv, hd = c.num_attention_heads, c.num_key_value_heads, c.hidden_size // c.num_attention_heads
        mask = torch.full((T, T), float("-inf"), dtype=self.dtype).triu(1)
        for i in range(c.num_hidden_layers):
            p = f"encoder.encoder_transformer.transformer.layers.{i}"
            n1 = F.layer_norm(h, (D,), self.w[f"{p}.norm1.weight"], self.w[f"{p}.norm1.bias"], 1e-5)
            q = (n1 @ self.w[f"{p}.self_attn.q_proj.weight"].T).reshape(B, T, nh, hd).transpose(1, 2)
            k = (n1 @ self.w[f"{p}.self_attn.k_proj.weight"].T).reshape(B, T, nkv, hd).transpose(1, 2)
            v = (n1 @ self.w[f"{p}.self_attn.v_proj.weight"].T).reshape(B, T, nkv, hd).transpose(1, 2)
            q, k = self.rope(q), self.rope(k)
            if nkv != nh:
                k = k.repeat_interleave(nh // nkv, dim=1)
                v = v.repeat_interleave(nh // nkv, dim=1)
            s = (q @ k.transpose(-1, -2)) * (hd ** -0.5) + mask
            a = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, T, nh * hd)
            h = h + self.w[f"{p}.layer_scale_1.scale"] * (a @ self.w[f"{p}.self_attn.o_proj.weight"].T)
            n2 = F.layer_norm(h, (D,), self.w[f"{p}.norm2.weight"], self.w[f"{p}.norm2.bias"], 1e-5)
            m = n2 @ self.w[f"{p}.gating.linear1.weight"].T
            m = m * 0.5 * (1.0 + torch.tanh(0.7978845608 * (m + 0.044715 * m ** 3)))     # geluApprox, STE.swift:1080-1082
            h = h + self.w[f"{p}.layer_scale_2.scale"] * (m @ self.w[f"{p}.gating.linear2.weight"].T)
            if taps is not None:
                taps[f"xf{i}"] = h.transpose(1, 2)
        return h.transpose(1, 2)

    def quantize(self, x, margins: Optional[list] = None):
        """x [B, C, T] -> codes [B, nvalid, T].  ``margins`` receives, per codebook, the gap between the two smallest distances [B, T]."""
        codes = []
        resid: Dict[str, torch.Tensor] = {}
        for part, i, E, c2 in self.books:
            if part not in resid:
                wproj = self.w[f"encoder.quantizer.{part}.input_proj.weight"][:, 0, :]       # [o, 1, i] MLX layout
                resid[part] = (x.transpose(1, 2) @ wproj.T)                                    # [B, T, cb_dim]
            r = resid[part]
            dist = c2.float() - r.float() @ E.float().T                                        # float32 by construction (STE.swift:750-757)
            idx = torch.argmin(dist, dim=-1)
            if margins is not None:
                two = torch.topk(dist, 2, dim=-1, largest=False).values
                margins.append((two[..., 1] - two[..., 0]))
            resid[part] = (r.float() - E.float()[idx]).to(self.dtype)
            codes.append(idx)
        return torch.stack(codes, dim=1)

    @torch.no_grad()
    def encode(self, audio, taps: Optional[dict] = None, margins: Optional[list] = None) -> torch.Tensor:
        """audio [B, 1, samples] float -> codes [B, 16, T] int64 (STE.swift:1031-1056)."""
        x = torch.as_tensor(np.asarray(audio), dtype=self.dtype)
        x = self.seanet(x, taps)
        x = self.transformer(x, taps)
        if taps is not None:
            taps["transformer"] = x
        s = self.cfg.downsample_stride
        x = self.sconv(x, "encoder.downsample.conv", 2 * s, stride=s, bias=False)
        if taps is not None:
            taps["downsample"] = x
        return self.quantize(x, margins)
