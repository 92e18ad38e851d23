"""Seeded synthetic checkpoint generator for the speech-tokenizer decoder.

Writes ``<dir>/speech_tokenizer/{config.json, model.safetensors}`` in the ON-DISK
(PyTorch / HF) layout the reference loads (Qwen3.swift:1461-1494): numeric
``decoder.decoder.N`` keys, ``[out,in,k]`` conv weights, ``[in,out,k]`` transposed-conv
weights and ``_codebook.{cluster_usage,embedding_sum}`` pairs (SURVEY Appendix B,
271 tensors).  The same file feeds the oracle and the CUDA engine.

Swift's default initialisers give a degenerate decoder (zero projections,
SpeechTokenizer.swift:107, 242-243, 382 -- SURVEY F6), so the "random-init weights of
the named architecture" are drawn here.  Gains are chosen so that every stage keeps
O(1) activations and the PCM is well inside [-1, 1] (clipping would hide errors).
FIXTURE TOOLING (see tools/__init__.py).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from .q3cfg import DecoderConfig, EncoderConfig, TokenizerConfig

DEFAULT_SEED = 20261018
# Global scale of outConv, calibrated once with tools in tests (PCM std ~0.15).
OUT_CONV_GAIN = 0.045


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _normal(g, shape, std, mean=0.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def decoder_tensor_specs(cfg: DecoderConfig):
    """(on-disk key, shape, kind, fan_in) in a fixed order; kind selects the initialiser."""
    specs = []
    half = cfg.codebook_dim // 2
    q = "decoder.quantizer"
    # codebooks: Qwen3.swift:1716-1724, SpeechTokenizer.swift:139, 196-208
    for i in range(cfg.num_semantic_quantizers):
        specs.append((f"{q}.rvq_first.vq.layers.{i}._codebook.cluster_usage", (cfg.semantic_codebook_size,), "usage", 0))
        specs.append((f"{q}.rvq_first.vq.layers.{i}._codebook.embedding_sum", (cfg.semantic_codebook_size, half), "embsum", 0))
    for i in range(cfg.num_quantizers - cfg.num_semantic_quantizers):
        specs.append((f"{q}.rvq_rest.vq.layers.{i}._codebook.cluster_usage", (cfg.codebook_size,), "usage", 0))
        specs.append((f"{q}.rvq_rest.vq.layers.{i}._codebook.embedding_sum", (cfg.codebook_size, half), "embsum", 0))
    for part in ("rvq_first", "rvq_rest"):
        specs.append((f"{q}.{part}.input_proj.weight", (half, cfg.codebook_dim, 1), "w", cfg.codebook_dim))
        specs.append((f"{q}.{part}.output_proj.weight", (cfg.codebook_dim, half, 1), "w", half))
    specs.append(("decoder.pre_conv.conv.weight", (cfg.latent_dim, cfg.codebook_dim, 3), "w", cfg.codebook_dim * 3))
    specs.append(("decoder.pre_conv.conv.bias", (cfg.latent_dim,), "b", 0))
    H, L, I = cfg.hidden_size, cfg.latent_dim, cfg.intermediate_size
    A = cfg.num_attention_heads * cfg.head_dim
    KV = cfg.num_key_value_heads * cfg.head_dim
    pt = "decoder.pre_transformer"
    specs += [(f"{pt}.input_proj.weight", (H, L), "w", L), (f"{pt}.input_proj.bias", (H,), "b", 0),
              (f"{pt}.output_proj.weight", (L, H), "w", H), (f"{pt}.output_proj.bias", (L,), "b", 0),
              (f"{pt}.norm.weight", (H,), "norm", 0)]
    for n in range(cfg.num_hidden_layers):
        p = f"{pt}.layers.{n}"
        specs += [(f"{p}.self_attn.q_proj.weight", (A, H), "w", H),
                  (f"{p}.self_attn.k_proj.weight", (KV, H), "w", H),
                  (f"{p}.self_attn.v_proj.weight", (KV, H), "w", H),
                  (f"{p}.self_attn.o_proj.weight", (H, A), "w", A),
                  (f"{p}.mlp.gate_proj.weight", (I, H), "w", H),
                  (f"{p}.mlp.up_proj.weight", (I, H), "w", H),
                  (f"{p}.mlp.down_proj.weight", (H, I), "w", I),
                  (f"{p}.input_layernorm.weight", (H,), "norm", 0),
                  (f"{p}.post_attention_layernorm.weight", (H,), "norm", 0),
                  (f"{p}.self_attn_layer_scale.scale", (H,), "lscale", 0),
                  (f"{p}.mlp_layer_scale.scale", (H,), "lscale", 0)]
    for i, r in enumerate(cfg.upsampling_ratios):
        u = f"decoder.upsample.{i}"
        specs += [(f"{u}.0.conv.weight", (L, L, r), "wt", L), (f"{u}.0.conv.bias", (L,), "b", 0),
                  (f"{u}.1.dwconv.conv.weight", (L, 1, 7), "w", 7), (f"{u}.1.dwconv.conv.bias", (L,), "b", 0),
                  (f"{u}.1.norm.weight", (L,), "norm", 0), (f"{u}.1.norm.bias", (L,), "b", 0),
                  (f"{u}.1.pwconv1.weight", (4 * L, L), "w", L), (f"{u}.1.pwconv1.bias", (4 * L,), "b", 0),
                  (f"{u}.1.pwconv2.weight", (L, 4 * L), "w", 4 * L), (f"{u}.1.pwconv2.bias", (L,), "b", 0),
                  (f"{u}.1.gamma", (L,), "lscale", 0)]
    D = cfg.decoder_dim
    dd = "decoder.decoder"
    specs += [(f"{dd}.0.conv.weight", (D, L, 7), "w", L * 7), (f"{dd}.0.conv.bias", (D,), "b", 0)]
    for i, r in enumerate(cfg.upsample_rates):
        cin, cout = D >> i, D >> (i + 1)
        b = f"{dd}.{i + 1}.block"
        specs += [(f"{b}.0.alpha", (cin,), "snake", 0), (f"{b}.0.beta", (cin,), "snake", 0),
                  # transposed conv: every output sample sees 2 taps x cin inputs
                  (f"{b}.1.conv.weight", (cin, cout, 2 * r), "wt", 2 * cin), (f"{b}.1.conv.bias", (cout,), "b", 0)]
        for j in (2, 3, 4):
            specs += [(f"{b}.{j}.act1.alpha", (cout,), "snake", 0), (f"{b}.{j}.act1.beta", (cout,), "snake", 0),
                      (f"{b}.{j}.conv1.conv.weight", (cout, cout, 7), "wres", cout * 7),
                      (f"{b}.{j}.conv1.conv.bias", (cout,), "b", 0),
                      (f"{b}.{j}.act2.alpha", (cout,), "snake", 0), (f"{b}.{j}.act2.beta", (cout,), "snake", 0),
                      (f"{b}.{j}.conv2.conv.weight", (cout, cout, 1), "wres", cout),
                      (f"{b}.{j}.conv2.conv.bias", (cout,), "b", 0)]
    cl = D >> len(cfg.upsample_rates)
    specs += [(f"{dd}.5.alpha", (cl,), "snake", 0), (f"{dd}.5.beta", (cl,), "snake", 0),
              (f"{dd}.6.conv.weight", (1, cl, 7), "wout", cl * 7), (f"{dd}.6.conv.bias", (1,), "b0", 0)]
    return specs


def make_decoder_state(cfg: DecoderConfig, seed: int = DEFAULT_SEED, out_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    state: Dict[str, torch.Tensor] = {}
    for key, shape, kind, fan_in in decoder_tensor_specs(cfg):
        if kind == "usage":
            t = torch.rand(shape, generator=g, dtype=torch.float32) * 99.0 + 1.0
            state[key] = t
        elif kind == "embsum":
            usage = state[key.replace("embedding_sum", "cluster_usage")]
            t = _normal(g, shape, 1.0) * usage[:, None]
        elif kind == "w":      # main path: unit gain, var = 1/fan_in
            t = _uniform(g, shape, (3.0 / fan_in) ** 0.5)
        elif kind == "wt":     # transposed conv (on-disk [in,out,k]), unit gain
            t = _uniform(g, shape, (3.0 / fan_in) ** 0.5)
        elif kind == "wres":   # residual-branch convs: MLXNN's own U(-1/sqrt(fan_in), ..) family
            t = _uniform(g, shape, (1.0 / fan_in) ** 0.5)
        elif kind == "wout":
            t = _uniform(g, shape, (3.0 / fan_in) ** 0.5) * (OUT_CONV_GAIN * out_gain)
        elif kind == "b":
            t = _normal(g, shape, 0.02)
        elif kind == "b0":
            t = torch.zeros(shape, dtype=torch.float32)
        elif kind == "snake":  # log-domain; e^a ~ 0.7..1.4 (pretrained means 0.82 / 0.96, Tests.swift:187)
            t = _normal(g, shape, 0.3)
        elif kind == "norm":
            t = _normal(g, shape, 0.1, mean=1.0)
        elif kind == "lscale":  # layer-scale / gamma: large enough that the branches matter
            t = torch.rand(shape, generator=g, dtype=torch.float32) * 0.45 + 0.05
        else:
            raise AssertionError(kind)
        state[key] = t.contiguous()
    return state


def encoder_tensor_specs(cfg: EncoderConfig):
    """(on-disk key, shape, kind, fan_in) of the speech-tokenizer ENCODER in the PyTorch / HF layout the reference remaps
    (Qwen3.swift:1514-1527: ``encoder.encoder.layers.N``; 1627-1649 transformer names; 1661-1676 quantizer names)."""
    specs = []
    e = "encoder.encoder.layers"
    nf, mult = cfg.num_filters, 1
    specs += [(f"{e}.0.conv.weight", (nf, cfg.audio_channels, cfg.kernel_size), "w", cfg.audio_channels * cfg.kernel_size),
              (f"{e}.0.conv.bias", (nf,), "b", 0)]
    idx = 1
    for ratio in reversed(cfg.upsampling_ratios):        # python layers: [res, elu, down] per ratio -> indices 1,3 / 4,6 / ...
        dim = mult * nf
        hid = dim // cfg.compress
        specs += [(f"{e}.{idx}.block.1.conv.weight", (hid, dim, cfg.residual_kernel_size), "wres", dim * cfg.residual_kernel_size),
                  (f"{e}.{idx}.block.1.conv.bias", (hid,), "b", 0),
                  (f"{e}.{idx}.block.3.conv.weight", (dim, hid, 1), "wres", hid),
                  (f"{e}.{idx}.block.3.conv.bias", (dim,), "b", 0),
                  (f"{e}.{idx + 2}.conv.weight", (2 * dim, dim, 2 * ratio), "w", dim * 2 * ratio),
                  (f"{e}.{idx + 2}.conv.bias", (2 * dim,), "b", 0)]
        idx += 3
        mult *= 2
    specs += [(f"{e}.{idx + 1}.conv.weight", (cfg.hidden_size, mult * nf, cfg.last_kernel_size), "w", mult * nf * cfg.last_kernel_size),
              (f"{e}.{idx + 1}.conv.bias", (cfg.hidden_size,), "b", 0)]
    H, I = cfg.hidden_size, cfg.intermediate_size
    hd = H // cfg.num_attention_heads
    KV = cfg.num_key_value_heads * hd
    for n in range(cfg.num_hidden_layers):
        p = f"encoder.encoder_transformer.layers.{n}"
        specs += [(f"{p}.input_layernorm.weight", (H,), "norm", 0), (f"{p}.input_layernorm.bias", (H,), "b", 0),
                  (f"{p}.post_attention_layernorm.weight", (H,), "norm", 0), (f"{p}.post_attention_layernorm.bias", (H,), "b", 0),
                  (f"{p}.self_attn.q_proj.weight", (H, H), "w", H), (f"{p}.self_attn.k_proj.weight", (KV, H), "w", H),
                  (f"{p}.self_attn.v_proj.weight", (KV, H), "w", H), (f"{p}.self_attn.o_proj.weight", (H, H), "w", H),
                  (f"{p}.mlp.fc1.weight", (I, H), "w", H), (f"{p}.mlp.fc2.weight", (H, I), "w", I),
                  (f"{p}.self_attn_layer_scale.scale", (H,), "lscale", 0), (f"{p}.mlp_layer_scale.scale", (H,), "lscale", 0)]
    s = cfg.downsample_stride
    specs.append(("encoder.downsample.conv.weight", (H, H, 2 * s), "w", H * 2 * s))
    q = "encoder.quantizer"
    for part, n in (("semantic_residual_vector_quantizer", 1), ("acoustic_residual_vector_quantizer", cfg.num_quantizers - 1)):
        specs += [(f"{q}.{part}.input_proj.weight", (cfg.codebook_dim, H, 1), "w", H),
                  (f"{q}.{part}.output_proj.weight", (H, cfg.codebook_dim, 1), "w", cfg.codebook_dim)]
        for i in range(n):
            specs += [(f"{q}.{part}.layers.{i}.codebook.cluster_usage", (cfg.codebook_size,), "usage", 0),
                      (f"{q}.{part}.layers.{i}.codebook.embed_sum", (cfg.codebook_size, cfg.codebook_dim), "embsum_enc", 0),
                      (f"{q}.{part}.layers.{i}.codebook.initialized", (1,), "flag", 0)]
    return specs


def make_encoder_state(cfg: EncoderConfig, seed: int = DEFAULT_SEED) -> Dict[str, torch.Tensor]:
    """Random-init encoder weights of the named architecture (the Swift initialisers give zero biases / tiny uniform weights,
    STE.swift:243-254; drawn here so that every stage carries signal).  Codebook l of a residual quantizer is drawn at the scale
    of the residual that reaches it (x 0.8 per layer), so that nearest-neighbour search stays meaningful down the chain."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 77)
    state: Dict[str, torch.Tensor] = {}
    for key, shape, kind, fan_in in encoder_tensor_specs(cfg):
        if kind == "usage":
            t = torch.rand(shape, generator=g, dtype=torch.float32) * 99.0 + 1.0
        elif kind == "embsum_enc":
            usage = state[key.replace("embed_sum", "cluster_usage")]
            layer = int(key.split(".layers.")[1].split(".")[0])
            t = _normal(g, shape, 0.6 * (0.8 ** layer)) * usage[:, None]
        elif kind == "flag":
            t = torch.ones(shape, dtype=torch.float32)
        elif kind == "w":
            t = _uniform(g, shape, (3.0 / fan_in) ** 0.5)
        elif kind == "wres":
            t = _uniform(g, shape, (1.0 / fan_in) ** 0.5)
        elif kind == "b":
            t = _normal(g, shape, 0.02)
        elif kind == "norm":
            t = _normal(g, shape, 0.1, mean=1.0)
        elif kind == "lscale":
            t = torch.rand(shape, generator=g, dtype=torch.float32) * 0.45 + 0.05
        else:
            raise AssertionError(kind)
        state[key] = t.contiguous()
    return state


def synth_audio(B: int, samples: int, seed: int) -> np.ndarray:
    """[B, 1, samples] float32: a few decaying sinusoids plus noise, |x| < 1 (speech-like dynamics, deterministic)."""
    rng = np.random.default_rng(seed)
    t = np.arange(samples, dtype=np.float64) / 24000.0
    out = np.zeros((B, 1, samples), dtype=np.float64)
    for b in range(B):
        for _ in range(6):
            f = rng.uniform(80.0, 4000.0)
            out[b, 0] += rng.uniform(0.02, 0.15) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) * (0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(1, 8) * t))
        out[b, 0] += rng.normal(0, 0.02, size=samples)
    return np.clip(out, -1.0, 1.0).astype(np.float32)


def write_checkpoint(model_dir: str, cfg: Optional[DecoderConfig] = None, seed: int = DEFAULT_SEED,
                     dtype: str = "float32", with_encoder_stub: bool = False,
                     mlx_layout: bool = False, out_gain: float = 1.0, encoder_cfg: Optional[EncoderConfig] = None) -> str:
    """Write ``<model_dir>/speech_tokenizer/`` and return that path.

    dtype 'float16' + no encoder == the 'lite' variant's on-disk form.
    ``with_encoder_stub`` adds a couple of ``encoder.*`` tensors + ``encoder_config`` that a
    decoder-only loader must ignore.  ``mlx_layout`` stores conv / transposed-conv weights
    pre-transposed to exercise the layout heuristic's "already MLX" branch.  ``out_gain`` scales outConv
    (x8 drives a good part of the PCM past [-1, 1] so that the final clip, ST.swift:781, is exercised).
    """
    from safetensors.torch import save_file
    cfg = cfg or DecoderConfig()
    st_dir = os.path.join(model_dir, "speech_tokenizer")
    os.makedirs(st_dir, exist_ok=True)
    state = make_decoder_state(cfg, seed, out_gain)
    if mlx_layout:
        for k in list(state.keys()):
            v = state[k]
            if v.ndim != 3 or "quantizer" in k:
                continue  # the quantizer projections are transposed unconditionally by the reference
            is_t = (".0.conv.weight" in k and "upsample" in k) or (".block.1.conv.weight" in k)
            state[k] = (v.permute(1, 2, 0) if is_t else v.permute(0, 2, 1)).contiguous()
    tdtype = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[dtype]
    out = {k: v.to(tdtype).contiguous() for k, v in state.items()}
    # decode_upsample_rate is a separate tokenizer-level key (Cfg.swift:590); keep it consistent with the decoder
    tok = TokenizerConfig(decoder_config=cfg, decode_upsample_rate=cfg.total_upsample, encode_downsample_rate=cfg.total_upsample)
    if with_encoder_stub:
        out["encoder.encoder.layers.0.conv.weight"] = torch.zeros(4, 1, 7, dtype=tdtype)
        out["encoder.quantizer.semantic_residual_vector_quantizer.layers.0.codebook.embed_sum"] = torch.zeros(8, 4, dtype=tdtype)
        tok.encoder_config = {"num_filters": 4}
    if encoder_cfg is not None:   # a real encoder, in a second shard (the reference merges every *.safetensors of the directory)
        enc = {k: v.to(tdtype).contiguous() for k, v in make_encoder_state(encoder_cfg, seed).items()}
        save_file(enc, os.path.join(st_dir, "model-encoder.safetensors"))
        tok.encoder_config = encoder_cfg.to_dict()
    save_file(out, os.path.join(st_dir, "model.safetensors"))
    with open(os.path.join(st_dir, "config.json"), "w") as f:
        json.dump(tok.to_dict(), f, indent=1)
    return st_dir


def synth_codes(cfg: DecoderConfig, B: int, T: int, seed: int, zero_frac: float = 0.0) -> np.ndarray:
    """Codes [B, 16, T] int32: codebook 0 ~ U{1..codebook_size-1} (what ``generate`` can emit,
    Qwen3.swift:622-628), codebooks 1.. ~ U{0..codebook_size-1}."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, cfg.codebook_size, size=(B, cfg.num_quantizers, T), dtype=np.int64)
    codes[:, 0, :] = rng.integers(1, cfg.codebook_size, size=(B, T), dtype=np.int64)
    if zero_frac > 0:
        mask = rng.random((B, T)) < zero_frac
        codes[:, 0, :][mask] = 0
    return codes.astype(np.int32)


def write_codec_embeddings(model_dir: str, hidden: int = 2048, talker_vocab: int = 3072, vocab: int = 2048, groups: int = 16,
                           dtype: str = "bfloat16", seed: int = DEFAULT_SEED) -> str:
    """Write ``<model_dir>/model.safetensors`` holding the codec-embedding tables of the MAIN checkpoint under the
    reference's key names (Talker.swift:495, CodePredictor.swift:206) plus two unrelated tensors a loader must skip.
    Values ~ N(0, 0.5): wide enough that the intermediate roundings of the 16-bit sequential sum matter."""
    from safetensors.torch import save_file
    os.makedirs(model_dir, exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    tdtype = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[dtype]
    out = {"talker.model.codec_embedding.weight": (torch.randn(talker_vocab, hidden, generator=g) * 0.5).to(tdtype)}
    for i in range(groups - 1):
        out[f"talker.code_predictor.model.codec_embedding.{i}.weight"] = (torch.randn(vocab, hidden, generator=g) * 0.5).to(tdtype)
    out["talker.model.text_embedding.weight"] = torch.zeros(4, 8, dtype=tdtype)
    out["talker.model.norm.weight"] = torch.ones(hidden, dtype=tdtype)
    save_file(out, os.path.join(model_dir, "model.safetensors"))
    return model_dir

