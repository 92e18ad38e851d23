"""CPU oracle for the 12 Hz speech-tokenizer decoder path (codes -> 24 kHz PCM).

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a CPU restatement (torch-CPU / NumPy)
of the reference decoder in
``/root/reference/Sources/Qwen3TTS/Models/SpeechTokenizer.swift:18-836`` and of the
weight-sanitize rules in ``.../Models/Qwen3.swift:1246-1260, 1498-1724``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the
CPU baseline.  The product (``libqwen3tts_cuda.so`` and the ``qwen3tts_cuda``
host wrapper) never imports, links or executes anything from here and has no
CPU fallback.

PARITY UNPINNED at the MLX boundary: the reference's arithmetic lives in the
un-vendored dependency ``mlx-swift`` 0.29.1 (rev 072b684a..., Package.resolved:13-20),
there is no Swift toolchain and no MLX build in this image, and the only numeric
assertions the reference's tests make on this path need the pretrained checkpoint
(Tests/Qwen3TTSTests/Qwen3TTSTests.swift:274-275).  What *is* pinned here, against
the reference's own fixtures: the 5x16 golden code grid (Tests.swift:37-43), the
per-stage shape chain (Tests.swift:69-253), the initConv weight layout KAT
(Tests.swift:131-132), the 271-tensor / 114.3 M-parameter inventory
(docs/paper.tex:218, 554), and the layout heuristic table (Qwen3.swift:1246-1260).
Op semantics follow MLX's documented definitions, cross-checked by a second,
definition-level implementation (``ops_def.py``) that shares no code with the
fast torch path.

``encoder.py`` is the same kind of restatement for the inverse path (SURVEY 8(f) row N3,
``SpeechTokenizerEncoder.swift:955-1056``): Seanet encoder, RoPE transformer, downsample,
split residual vector quantizer; also parity-unpinned at the MLX boundary (the reference
has no test or golden vector for ``encode``), pinned against tap-by-tap definitions,
the padding rule's known answers and the key-remap table in ``tests/test_oracle_encoder.py``.
"""

from tools.q3cfg import DecoderConfig, EncoderConfig, TokenizerConfig, GOLDEN_CODES_5x16  # noqa: F401  (hyper-parameters live in tools/q3cfg.py so that bench.py can build a synthetic checkpoint without importing the oracle)
