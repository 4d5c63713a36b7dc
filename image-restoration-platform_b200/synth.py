"""Synthetic inputs of BASELINE.md §4 (no datasets offline).

u8 interleaved RGB: v = clip(128 + 64 sin(6 pi x / W) cos(4 pi y / H) + N(0, 12^2) per channel),
blended 75/25 with its 8x8 block mean so stencil, scratch and block-edge counts are non-trivial.
Seed = 0xB200 + idx (numpy PCG64).
"""
from __future__ import annotations

import numpy as np


def synth_image(width: int, height: int, idx: int = 0, channels: int = 3) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(0xB200 + idx))
    x = np.arange(width, dtype=np.float32)[None, :]
    y = np.arange(height, dtype=np.float32)[:, None]
    base = 128.0 + 64.0 * np.sin(6.0 * np.pi * x / width) * np.cos(4.0 * np.pi * y / height)
    img = base[:, :, None] + rng.normal(0.0, 12.0, (height, width, channels)).astype(np.float32)
    hb, wb = height // 8 * 8, width // 8 * 8
    if hb and wb:
        blk = img[:hb, :wb].reshape(hb // 8, 8, wb // 8, 8, channels).mean(axis=(1, 3), keepdims=True)
        blk = np.broadcast_to(blk, (hb // 8, 8, wb // 8, 8, channels)).reshape(hb, wb, channels)
        img[:hb, :wb] = 0.75 * img[:hb, :wb] + 0.25 * blk
    # a few bright hairlines so the scratch detector has something to count
    for k in range(3):
        col = int(rng.integers(0, width))
        img[:, col : col + 1] += 180.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_batch(width: int, height: int, n: int, distinct: int = 8, channels: int = 3):
    """n images of which `distinct` are generated and the rest are row-rolled copies (cheap, still
    different content per image)."""
    bases = [synth_image(width, height, i, channels) for i in range(min(n, distinct))]
    out = []
    for i in range(n):
        b = bases[i % len(bases)]
        out.append(b if i < len(bases) else np.roll(b, 37 * (i // len(bases)), axis=0))
    return out


def mixed_resolution_sizes(n: int = 512, seed: int = 0xB200C5, mp_lo: float = 0.5, mp_hi: float = 24.0):
    """BASELINE.json configs[4]: MP log-uniform in [0.5, 24], aspect from a fixed set."""
    rng = np.random.Generator(np.random.PCG64(seed))
    aspects = [(4, 3), (3, 2), (16, 9), (1, 1), (3, 4), (2, 3)]
    sizes = []
    for _ in range(n):
        mp = float(np.exp(rng.uniform(np.log(mp_lo), np.log(mp_hi)))) * 1e6
        a, b = aspects[int(rng.integers(0, len(aspects)))]
        w = int(round((mp * a / b) ** 0.5))
        h = int(round(w * b / a))
        sizes.append((max(w, 16), max(h, 16)))
    return sizes
