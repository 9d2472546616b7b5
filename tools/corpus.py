"""Seeded synthetic mixed-entropy corpus.

One PCG64 generator drives everything, so a (seed, size, mix) triple is reproducible
anywhere.  A corpus is a concatenation of segments drawn with probabilities
``mix = (p_random, p_words, p_runs)``:

* random : U[64, 65536] uniform random bytes            (incompressible)
* words  : U[100, 20000] words drawn from a 4096-word dictionary of 3..12-byte random words
* runs   : one random byte repeated U[16, 8192] times   (overlapping rep0 matches)

``x86_callsites`` sprinkles ``E8 rel32`` call sites (high byte 00/FF) so the BCJ filter
has work to do (config 4).
"""
from __future__ import annotations

import numpy as np

DEFAULT_MIX = (0.25, 0.50, 0.25)


def make_corpus(size: int, seed: int = 1234, mix=DEFAULT_MIX) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    n_words = 4096
    wlen = rng.integers(3, 13, size=n_words)
    wstart = np.concatenate(([0], np.cumsum(wlen)))
    wbytes = rng.integers(0, 256, size=int(wstart[-1]), dtype=np.uint8)
    out = np.empty(size, dtype=np.uint8)
    pos = 0
    p = np.asarray(mix, dtype=np.float64)
    p = p / p.sum()
    while pos < size:
        kind = int(rng.choice(3, p=p))
        if kind == 0:
            n = int(rng.integers(64, 65537))
            seg = rng.integers(0, 256, size=n, dtype=np.uint8)
        elif kind == 1:
            k = int(rng.integers(100, 20001))
            ids = rng.integers(0, n_words, size=k)
            lens = wlen[ids]
            total = int(lens.sum())
            # gather variable-length words without a python loop
            starts = wstart[ids]
            offs = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens)
            seg = wbytes[np.arange(total) + offs]
        else:
            n = int(rng.integers(16, 8193))
            seg = np.full(n, int(rng.integers(0, 256)), dtype=np.uint8)
        n = min(len(seg), size - pos)
        out[pos:pos + n] = seg[:n]
        pos += n
    return out


def add_x86_callsites(buf: np.ndarray, seed: int = 99, every: int = 97) -> np.ndarray:
    """Overwrite ~1/every of the positions with ``E8/E9 lo lo lo 00|FF`` (little-endian rel32)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = buf.copy()
    n = len(out)
    if n < 16:
        return out
    count = max(1, n // every)
    pos = np.sort(rng.integers(0, n - 5, size=count))
    # also create adjacent/overlapping candidates to exercise the prevMask logic
    extra = pos[: count // 8] + rng.integers(1, 4, size=count // 8)
    pos = np.unique(np.clip(np.concatenate((pos, extra)), 0, n - 6))
    out[pos] = np.where(rng.integers(0, 4, size=len(pos)) == 0, 0xE9, 0xE8).astype(np.uint8)
    out[pos + 4] = np.where(rng.integers(0, 2, size=len(pos)) == 0, 0x00, 0xFF).astype(np.uint8)
    return out


def entropy_sweep_mix(t: float):
    """t=0 -> all runs (lowest entropy), t=1 -> all random; word-soup peaks in the middle."""
    t = float(min(max(t, 0.0), 1.0))
    return (t * t, 2.0 * t * (1.0 - t) + 1e-9, (1.0 - t) * (1.0 - t))
