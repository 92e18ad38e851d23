"""Comparison of two code grids produced by a residual vector quantizer (tests only).  A code is an argmin over float32 distances:
two correct implementations that sum in different orders may disagree where the two best distances are a rounding error apart,
and from that codebook on the frame's residual -- hence every later code of the frame -- legitimately differs."""
from __future__ import annotations

import numpy as np


def count_near_tie_frames(want: np.ndarray, got: np.ndarray, margins, tol: float) -> int:
    """want / got: [B, Q, T]; margins[q]: [B, T] gap between the two smallest distances of codebook q on the reference path.
    Asserts that a frame leaves the reference path only at a codebook whose gap is below ``tol``; returns how many frames did."""
    assert want.shape == got.shape, (want.shape, got.shape)
    B, Q, T = want.shape
    bad = 0
    for b in range(B):
        for t in range(T):
            for qi in range(Q):
                if want[b, qi, t] != got[b, qi, t]:
                    gap = float(np.asarray(margins[qi])[b, t])
                    assert gap < tol, f"utterance {b} frame {t} codebook {qi}: {want[b, qi, t]} vs {got[b, qi, t]}, gap {gap:.3e}"
                    bad += 1
                    break
    return bad
