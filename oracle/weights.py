"""Restatement of the reference's speech-tokenizer weight sanitize rules (decoder keys).

Follows Sources/Qwen3TTS/Models/Qwen3.swift:
  1246-1260  checkArrayShapeQwen3 (layout heuristic)
  1504-1512  decoder.decoder.{0..6} -> initConv, block0..3, outSnake, outConv
  1581-1588  .block.{0..4}. -> snake, upsample, res1..3
  1532-1543 + 1716-1724  *._codebook.{cluster_usage,embedding_sum} -> codebook.embed.weight
  1688-1692  quantizer input_proj/output_proj [o,i,1] -> [o,1,i] (unconditional)
  1696-1700  conv weights [o,i,k] -> [o,k,i] unless the heuristic says "already MLX"
  1704-1711  transposed-conv weights [i,o,k] -> [o,k,i] (overrides the generic branch)
The result is a dict keyed by the Swift module path with arrays in MLX layout
(conv weight [Cout, K, Cin]).  TEST INFRASTRUCTURE (see package docstring).
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

DECODER_INDEX_MAPPING = {
    "decoder.decoder.0": "decoder.decoder.initConv",
    "decoder.decoder.1": "decoder.decoder.block0",
    "decoder.decoder.2": "decoder.decoder.block1",
    "decoder.decoder.3": "decoder.decoder.block2",
    "decoder.decoder.4": "decoder.decoder.block3",
    "decoder.decoder.5": "decoder.decoder.outSnake",
    "decoder.decoder.6": "decoder.decoder.outConv",
}


def is_mlx_conv_layout(shape) -> bool:
    """Qwen3.swift:1246-1260: True => leave as is, False => transpose."""
    if len(shape) != 3:
        return False
    _, d2, d3 = shape
    if d2 == 1:
        return d3 > 64
    if d3 == 1:
        return d2 <= 64
    return d2 < d3


def sanitize_decoder_weights(weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Qwen3.swift:1498-1750 restricted to ``decoder.*`` keys (encoder.* are dropped)."""
    out: Dict[str, np.ndarray] = {}
    codebooks: Dict[str, Dict[str, np.ndarray]] = {}
    for key, value in weights.items():
        if key.startswith("encoder."):
            continue  # out of scope (the 'lite' variant has none)
        if "._codebook.cluster_usage" in key or "._codebook.embedding_sum" in key:
            base = key.split("._codebook.")[0]
            slot = "cluster_usage" if "cluster_usage" in key else "embedding_sum"
            codebooks.setdefault(base, {})[slot] = value
            continue
        new_key = key
        for idx_prefix, named in DECODER_INDEX_MAPPING.items():
            if key.startswith(idx_prefix):
                new_key = key.replace(idx_prefix, named)
                break
        if new_key.startswith("decoder."):
            new_key = (new_key.replace(".block.0.", ".snake.")
                       .replace(".block.1.", ".upsample.")
                       .replace(".block.2.", ".res1.")
                       .replace(".block.3.", ".res2.")
                       .replace(".block.4.", ".res3."))
        new_value = value
        is_proj = (("input_proj.weight" in new_key or "output_proj.weight" in new_key)
                   and "quantizer" in new_key)
        if is_proj and value.ndim == 3:
            new_value = np.transpose(value, (0, 2, 1))
        if "conv.weight" in new_key and value.ndim == 3 and not is_proj:
            if not is_mlx_conv_layout(value.shape):
                new_value = np.transpose(value, (0, 2, 1))
        is_tconv = (("upsample" in new_key and ".0.conv.weight" in new_key)
                    or ("decoder.decoder.block" in new_key and "upsample.conv.weight" in new_key))
        if is_tconv and value.ndim == 3:
            # NB: computed from the ORIGINAL value, overriding the generic branch (1704-1711)
            if not is_mlx_conv_layout(value.shape):
                new_value = np.transpose(value, (1, 2, 0))
            else:
                new_value = value
        out[new_key] = np.ascontiguousarray(new_value)
    eps = np.float32(1e-5)
    for base, data in codebooks.items():
        if "cluster_usage" in data and "embedding_sum" in data:
            usage = data["cluster_usage"].astype(np.float32)
            esum = data["embedding_sum"].astype(np.float32)
            denom = np.clip(usage[:, None], eps, np.finfo(np.float32).max)
            out[f"{base}.codebook.embed.weight"] = (esum / denom).astype(np.float32)
    return out


def load_safetensors_dir(speech_tokenizer_dir: str) -> Dict[str, np.ndarray]:
    """Merge every *.safetensors in the directory (Qwen3.swift:1473-1480), as float32."""
    from safetensors import safe_open
    import torch
    merged: Dict[str, np.ndarray] = {}
    for name in sorted(os.listdir(speech_tokenizer_dir)):
        if not name.endswith(".safetensors"):
            continue
        with safe_open(os.path.join(speech_tokenizer_dir, name), framework="pt") as f:
            for k in f.keys():
                merged[k] = f.get_tensor(k).to(torch.float32).numpy()
    return merged


def load_decoder(speech_tokenizer_dir: str):
    """config.json + weights -> (TokenizerConfig, sanitized MLX-layout weight dict)."""
    from tools.q3cfg import TokenizerConfig
    cfg = TokenizerConfig.from_json(os.path.join(speech_tokenizer_dir, "config.json"))
    if cfg.decoder_config is None:
        raise ValueError("Decoder config is required")  # SpeechTokenizer.swift:801-805
    raw = load_safetensors_dir(speech_tokenizer_dir)
    return cfg, sanitize_decoder_weights(raw)
