"""CPU oracle for SURVEY 8(f) row N2: the codec-embedding sum that builds the Talker's next-step input.

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, smoke() and bench.py's CPU legs may import it.

Restates, frame by frame,
    var codecEmbed = talker.getInputEmbeddings()(nextToken)
    for (i, code) in codeTokens.dropFirst().enumerated() { codecEmbed = codecEmbed + codePredictor.codecEmbedding[i](code) }
(/root/reference/Sources/Qwen3TTS/Models/Qwen3.swift:720-728, 927-935, 1157-1162) and its whole-prefix form for voice
cloning (Qwen3.swift:485-491): a left-to-right sequence of array adds in the checkpoint's dtype.  MLX evaluates a 16-bit
add in fp32 and rounds the result to the array dtype (round-to-nearest-even), which is what torch's CPU bf16 / fp16 add
does too, so the intermediate roundings are reproduced by adding torch tensors of that dtype in the same order.
PARITY UNPINNED at the MLX boundary (no MLX build here); pinned against the float64 definition below within the rounding
bound, and bit-exactly against the CUDA kernel.
"""
from typing import List, Sequence

import numpy as np
import torch


def codec_embed_sum(tables: Sequence[torch.Tensor], codes: np.ndarray) -> torch.Tensor:
    """tables[0]: talker.model.codec_embedding.weight [V0,H]; tables[1+i]: code_predictor codec_embedding.i.weight [V,H].
    codes int [n, G] frame-major.  Returns [n, H] in the tables' dtype."""
    c = torch.as_tensor(np.asarray(codes), dtype=torch.long)
    assert c.ndim == 2 and c.shape[1] == len(tables)
    out = tables[0][c[:, 0]]                              # Embedding lookup = row gather (Talker.swift:617)
    for g in range(1, len(tables)):
        out = out + tables[g][c[:, g]]                    # one rounding to the dtype per add, in this order
    return out


def codec_embed_sum_f64(tables: Sequence[torch.Tensor], codes: np.ndarray) -> np.ndarray:
    """Definition-level check: exact sum in float64, no intermediate rounding."""
    c = np.asarray(codes, dtype=np.int64)
    acc = np.zeros((c.shape[0], tables[0].shape[1]), dtype=np.float64)
    for g, t in enumerate(tables):
        acc += t.to(torch.float64).numpy()[c[:, g]]
    return acc


def load_tables(model_dir: str) -> List[torch.Tensor]:
    """Read the tables from <model_dir>/*.safetensors under the reference's key names."""
    import glob, os
    from safetensors.torch import load_file
    st = {}
    for f in sorted(glob.glob(os.path.join(model_dir, "*.safetensors"))):
        st.update({k: v for k, v in load_file(f).items() if "codec_embedding" in k})
    tables = [st["talker.model.codec_embedding.weight"]]
    i = 0
    while f"talker.code_predictor.model.codec_embedding.{i}.weight" in st:
        tables.append(st[f"talker.code_predictor.model.codec_embedding.{i}.weight"])
        i += 1
    return tables
