"""Cached synthetic checkpoints on local disk (never committed: 459 MB for the full model)."""
from __future__ import annotations

import hashlib
import json
import os
import tempfile
from typing import Optional

from .q3cfg import DecoderConfig, EncoderConfig
from .synth_checkpoint import DEFAULT_SEED, write_checkpoint


def checkpoint_dir(cfg: Optional[DecoderConfig] = None, seed: int = DEFAULT_SEED, dtype: str = "float32",
                   with_encoder_stub: bool = False, mlx_layout: bool = False, root: Optional[str] = None,
                   out_gain: float = 1.0, encoder_cfg: Optional[EncoderConfig] = None) -> str:
    """Return ``<model_dir>`` containing ``speech_tokenizer/`` for this (config, seed, dtype); build once.
    ``encoder_cfg`` adds a real encoder (second safetensors shard + ``encoder_config``)."""
    cfg = cfg or DecoderConfig()
    key = json.dumps([cfg.to_dict(), seed, dtype, with_encoder_stub, mlx_layout] + ([out_gain] if out_gain != 1.0 else [])
                     + ([encoder_cfg.to_dict()] if encoder_cfg is not None else []), sort_keys=True)
    tag = hashlib.sha1(key.encode()).hexdigest()[:12]
    root = root or os.environ.get("Q3TTS_FIXTURE_ROOT") or os.path.join(tempfile.gettempdir(), "q3tts_fixtures")
    model_dir = os.path.join(root, f"ckpt_{tag}")
    done = os.path.join(model_dir, ".complete")
    if not os.path.exists(done):
        os.makedirs(model_dir, exist_ok=True)
        write_checkpoint(model_dir, cfg, seed, dtype, with_encoder_stub, mlx_layout, out_gain, encoder_cfg)
        with open(done, "w") as f:
            f.write(key)
    return model_dir
