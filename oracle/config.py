"""Re-export of the hyper-parameter dataclasses (they live in tools/q3cfg.py so that bench.py's
product arm can build a synthetic checkpoint without importing the oracle)."""
from tools.q3cfg import DecoderConfig, EncoderConfig, TokenizerConfig, GOLDEN_CODES_5x16  # noqa: F401
