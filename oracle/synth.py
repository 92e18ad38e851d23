"""Re-export of the synthetic checkpoint generator (tools/synth_checkpoint.py)."""
from tools.synth_checkpoint import *  # noqa: F401,F403
from tools.synth_checkpoint import DEFAULT_SEED, make_decoder_state, write_checkpoint, synth_codes, decoder_tensor_specs  # noqa: F401
