"""Regenerates the committed golden fixtures from the oracle (run here, on CPU):

    python tests/golden/make_golden.py

golden_full_5x16.npz  -- the reference's golden 5x16 code grid (Tests.swift:37-43) through the
                         full-size decoder with the seed-20261018 synthetic checkpoint, float64 oracle:
                         PCM (float32), audio_lengths, per-stage (std, min, max), the first 10 dequantised
                         values the reference's test prints (Tests.swift:73-75 pattern).
golden_tiny.npz       -- every stage tap of the tiny architecture (seed 7) for codes (B=1, T=6, seed 3), float64 oracle,
                         reference and causal_sw attention.
The reference itself cannot run in this image (Swift + MLX absent), so these pin the ORACLE, not MLX:
they catch oracle drift and let the GPU box check the CUDA path without /root/reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import decoder, weights  # noqa: E402
from tools.fixtures import checkpoint_dir  # noqa: E402
from tools.q3cfg import DecoderConfig, GOLDEN_CODES_5x16  # noqa: E402
from tools.synth_checkpoint import synth_codes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # full-size
    st = os.path.join(checkpoint_dir(DecoderConfig()), "speech_tokenizer")
    cfg, w = weights.load_decoder(st)
    codes = np.asarray(GOLDEN_CODES_5x16, dtype=np.int32)[None]
    taps = {}
    audio, lengths = decoder.OracleDecoder(cfg.decoder_config, w, torch.float64).decode(codes, taps)
    stats = {k: np.array([float(v.std()), float(v.min()), float(v.max())]) for k, v in taps.items()}
    np.savez_compressed(os.path.join(HERE, "golden_full_5x16.npz"), codes=codes, audio=audio.numpy().astype(np.float32),
                        lengths=lengths, quantized_0_10_0=taps["quantized"][0, :10, 0].numpy(),
                        **{f"stat_{k}": v for k, v in stats.items()},
                        **{f"shape_{k}": np.array(v.shape) for k, v in taps.items()})
    # tiny
    tcfg = DecoderConfig.tiny()
    st = os.path.join(checkpoint_dir(tcfg, seed=7), "speech_tokenizer")
    cfg, w = weights.load_decoder(st)
    tcodes = synth_codes(tcfg, 1, 6, 3)
    out = {"codes": tcodes}
    for mode in ("reference", "causal_sw"):
        taps = {}
        decoder.OracleDecoder(cfg.decoder_config, w, torch.float64, attn_mode=mode).forward(tcodes, taps)
        for k, v in taps.items():
            out[f"{mode}_{k}"] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "golden_tiny.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
