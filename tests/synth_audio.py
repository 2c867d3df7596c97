"""Bit-reproducible synthetic recordings for golden vectors that are too long to store: integer arithmetic on numpy's PCG64
stream only (no libm), so the script that makes a golden (scripts/make_glue_golden.py) and the test that replays it on another
machine see identical int16 samples."""
import numpy as np


def integer_meeting(seconds: float, seed: int, speakers: int = 4, sr: int = 16000) -> np.ndarray:
    """-> int16 (n,): `speakers` noise sources with different (box-filter) spectra taking turns of 2-12 s, some overlapping."""
    r = np.random.Generator(np.random.PCG64(seed))
    n = int(seconds * sr)
    out = np.zeros(n, dtype=np.int64)
    widths = (1, 3, 9, 27, 5, 15, 2, 45)
    pos = 0
    while pos < n:
        s = int(r.integers(0, speakers))
        length = int(r.integers(2 * sr, 12 * sr))
        start = max(0, pos - int(r.integers(0, sr)) * int(r.integers(0, 2)))      # every other turn starts up to 1 s early
        end = min(n, start + length)
        w = widths[s % len(widths)]
        noise = r.integers(-6000, 6001, size=end - start + w, dtype=np.int64)
        c = np.cumsum(noise)
        src = (c[w:] - c[:-w]) * 2 // (w + 1)                                    # box filter of width w, integer arithmetic
        out[start:end] += src[: end - start]
        pos = end + int(r.integers(0, sr // 2)) * int(r.integers(0, 2))
    return np.clip(out, -32768, 32767).astype(np.int16)
