"""Decoder / tokenizer hyper-parameters with the reference's defaults.

Follows Sources/Qwen3TTS/Models/Config.swift:338-415 (Qwen3TTSTokenizerDecoderConfig:
every field ``decodeIfPresent ?? default``, snake_case JSON keys at 361-383) and
Config.swift:565-595 (Qwen3TTSTokenizerConfig).  TEST INFRASTRUCTURE (see package
docstring).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import List, Optional

# Tests/Qwen3TTSTests/Qwen3TTSTests.swift:37-43 -- the reference's one golden input.
GOLDEN_CODES_5x16 = [
    [1342, 313, 975, 826, 309, 933, 1642, 28, 782, 1965, 1680, 1507, 258, 1349, 828, 1102],
    [1014, 17, 2016, 285, 1712, 470, 543, 176, 1087, 1963, 588, 1860, 889, 1874, 1121, 1319],
    [1119, 1122, 594, 89, 770, 1644, 331, 242, 1183, 1676, 129, 96, 123, 1840, 661, 285],
    [1119, 1135, 215, 1377, 88, 1611, 904, 1274, 1895, 1872, 1246, 335, 1141, 320, 694, 242],
    [46, 1952, 1023, 1871, 596, 491, 757, 422, 692, 683, 651, 395, 1235, 1300, 618, 1498],
]


@dataclass
class DecoderConfig:
    # Config.swift:388-408 defaults
    latent_dim: int = 1024
    codebook_dim: int = 512
    codebook_size: int = 2048
    decoder_dim: int = 1536
    hidden_size: int = 512
    intermediate_size: int = 1024
    num_hidden_layers: int = 8
    num_attention_heads: int = 16
    num_key_value_heads: int = 16
    head_dim: int = 64
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0           # parsed, unused by the decoder
    max_position_embeddings: int = 8000   # parsed, unused
    sliding_window: int = 72              # parsed, unused by the reference decoder
    num_quantizers: int = 16
    num_semantic_quantizers: int = 1
    semantic_codebook_size: int = 4096
    upsample_rates: List[int] = field(default_factory=lambda: [8, 5, 4, 3])
    upsampling_ratios: List[int] = field(default_factory=lambda: [2, 2])
    vector_quantization_hidden_dimension: int = 512  # parsed, unused
    layer_scale_initial_scale: float = 0.01

    @property
    def total_upsample(self) -> int:  # Config.swift:411-414
        r = 1
        for x in self.upsample_rates:
            r *= x
        for x in self.upsampling_ratios:
            r *= x
        return r

    @classmethod
    def from_dict(cls, d: dict) -> "DecoderConfig":
        known = {k: d[k] for k in cls.__dataclass_fields__ if k in d}
        return cls(**known)

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def tiny(cls) -> "DecoderConfig":
        """A scaled-down architecture (same topology) for fast CPU tests.

        decoder_dim stays >= 65*16 so that every 1x1 conv has > 64 channels: below that the
        reference's layout heuristic (Qwen3.swift:1246-1260, ``dim3 == 1 -> dim2 <= 64``)
        misreads a PyTorch ``[C, C, 1]`` weight as already-MLX."""
        return cls(latent_dim=80, codebook_dim=48, codebook_size=64, decoder_dim=1152,
                   hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                   num_attention_heads=4, num_key_value_heads=2, head_dim=32,
                   semantic_codebook_size=128, sliding_window=4,
                   upsample_rates=[3, 2, 2, 2], upsampling_ratios=[2, 2])


@dataclass
class EncoderConfig:
    """Qwen3TTSTokenizerEncoderConfig, Config.swift:419-560 (every field ``decodeIfPresent ?? default``)."""
    frame_rate: float = 12.5
    audio_channels: int = 1
    codebook_dim: int = 256
    codebook_size: int = 2048
    compress: int = 2
    dilation_growth_rate: int = 2
    head_dim: int = 64
    hidden_size: int = 512
    intermediate_size: int = 2048
    kernel_size: int = 7
    last_kernel_size: int = 3
    layer_scale_initial_scale: float = 0.01
    max_position_embeddings: int = 8000
    num_attention_heads: int = 8
    num_filters: int = 64
    num_hidden_layers: int = 8
    num_key_value_heads: int = 8
    num_quantizers: int = 32
    num_residual_layers: int = 1
    residual_kernel_size: int = 3
    rope_theta: float = 10000.0
    sampling_rate: int = 24000
    sliding_window: int = 250             # parsed; encode() builds a FULL causal mask (SpeechTokenizerEncoder.swift:1038-1042)
    upsampling_ratios: List[int] = field(default_factory=lambda: [8, 6, 5, 4])
    use_causal_conv: bool = True
    use_conv_shortcut: bool = False

    @property
    def seanet_stride(self) -> int:
        r = 1
        for x in self.upsampling_ratios:
            r *= x
        return r

    @property
    def downsample_stride(self) -> int:   # SpeechTokenizerEncoder.swift:1005-1006
        return int((self.sampling_rate / self.seanet_stride) / self.frame_rate)

    @classmethod
    def from_dict(cls, d: dict) -> "EncoderConfig":
        return cls(**{k: d[k] for k in cls.__dataclass_fields__ if k in d})

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def tiny(cls) -> "EncoderConfig":
        """Same topology, small widths, for CPU tests (total stride 2*2*3*2 * 2 = 48 samples per code frame)."""
        return cls(codebook_dim=16, codebook_size=32, head_dim=32, hidden_size=64, intermediate_size=128, num_attention_heads=2,
                   num_key_value_heads=2, num_filters=8, num_hidden_layers=2, num_quantizers=20, upsampling_ratios=[2, 3, 2, 2],
                   sampling_rate=24000, frame_rate=500.0)


@dataclass
class TokenizerConfig:
    # Config.swift:586-592 defaults
    encoder_valid_num_quantizers: int = 16
    input_sample_rate: int = 24000
    output_sample_rate: int = 24000
    decode_upsample_rate: int = 1920
    encode_downsample_rate: int = 1920
    decoder_config: Optional[DecoderConfig] = None
    encoder_config: Optional[dict] = None

    @classmethod
    def from_json(cls, path: str) -> "TokenizerConfig":
        with open(path) as f:
            d = json.load(f)
        out = cls()
        for k in ("encoder_valid_num_quantizers", "input_sample_rate", "output_sample_rate",
                  "decode_upsample_rate", "encode_downsample_rate"):
            if k in d:
                setattr(out, k, d[k])
        if d.get("decoder_config") is not None:
            out.decoder_config = DecoderConfig.from_dict(d["decoder_config"])
        out.encoder_config = d.get("encoder_config")
        return out

    def to_dict(self) -> dict:
        d = {k: getattr(self, k) for k in ("encoder_valid_num_quantizers", "input_sample_rate",
                                           "output_sample_rate", "decode_upsample_rate",
                                           "encode_downsample_rate")}
        if self.decoder_config is not None:
            d["decoder_config"] = self.decoder_config.to_dict()
        if self.encoder_config is not None:
            d["encoder_config"] = self.encoder_config
        return d
