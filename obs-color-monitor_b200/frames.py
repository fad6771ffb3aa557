"""Synthetic BGRA frames used by the tests and the benchmark (host side, numpy).

The definitions are the ones SURVEY.md §8(d) fixes for BASELINE.json's configs:

* ``ramp``    pixel(x, y) = (B = x & 255, G = y & 255, R = (x + y) & 255, A = 255)
* ``random``  ``numpy.random.default_rng(seed)`` uniform u8, alpha forced to 255
* ``solid``   one colour everywhere (worst case for bin contention / saturation)
* ``alpha_stripes``  random colours, every ``period``-th column fully transparent
  (exercises the ``a == 0`` skip of histogram.c:385-387 / waveform.c:246-248)
* ``natural`` smooth gradients + low-amplitude noise (video-like bin locality)
* ``ui``      screen-capture-like: flat background with text lines (almost-flat blocks; not part of ``mixed``)
"""
from __future__ import annotations

import numpy as np


def ramp(width: int, height: int) -> np.ndarray:
    x = np.arange(width, dtype=np.uint32)[None, :]
    y = np.arange(height, dtype=np.uint32)[:, None]
    f = np.empty((height, width, 4), np.uint8)
    f[..., 0] = (x & 255) + 0 * y
    f[..., 1] = (y & 255) + 0 * x
    f[..., 2] = (x + y) & 255
    f[..., 3] = 255
    return f


def random(width: int, height: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 256, size=(height, width, 4), dtype=np.uint8)
    f[..., 3] = 255
    return f


def solid(width: int, height: int, bgra=(128, 128, 128, 255)) -> np.ndarray:
    f = np.empty((height, width, 4), np.uint8)
    f[...] = np.asarray(bgra, np.uint8)
    return f


def alpha_stripes(width: int, height: int, seed: int = 1, period: int = 3) -> np.ndarray:
    f = random(width, height, seed)
    f[:, ::period, 3] = 0
    rng = np.random.default_rng(seed + 1000)
    # a sprinkle of partially transparent pixels: still counted (only a == 0 is skipped)
    m = rng.random((height, width)) < 0.05
    f[m & (f[..., 3] != 0), 3] = 7
    return f


def natural(width: int, height: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :]
    y = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None]
    f = np.empty((height, width, 4), np.uint8)
    base = [
        96 + 80 * np.sin(2.1 * x + 0.7 * y + seed),
        110 + 70 * np.cos(1.3 * x - 1.9 * y + 0.3 * seed),
        128 + 60 * np.sin(0.9 * x * y * 3.0 + 1.1),
    ]
    for c in range(3):
        n = rng.integers(-3, 4, size=(height, width), dtype=np.int16)
        f[..., c] = np.clip(base[c] + n, 0, 255).astype(np.uint8)
    f[..., 3] = 255
    return f


def ui(width: int, height: int, seed: int = 0) -> np.ndarray:
    """Screen-capture-like content (an editor window): a flat dark background, 12-pixel text lines every 24 rows
    with ~25 % ink in two colours.  Most 4 x 32 blocks are *almost* flat: nearly every lane of a warp hits the
    background's vectorscope bin, the worst case for same-bin serialisation that the mixed batch does not contain."""
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    f = np.empty((height, width, 4), np.uint8)
    f[...] = (34, 30, 30, 255)
    line = (y % 24 >= 6) & (y % 24 < 18)
    h = (((x // 2) * 73856093) ^ ((y // 2) * 19349663) ^ (seed * 83492791)) & 0xFFFFFFFF
    glyph = line & (((h >> 7) & 3) == 0) & (x % 512 < 400)
    word = ((((x // 64) * 2654435761) ^ ((y // 24) * 40503) ^ seed) & 0xFFFFFFFF) >> 11
    accent = (word % 5 == 0) & glyph
    plain = glyph & ~accent
    f[plain] = (220, 220, 220, 255)
    f[accent] = (90, 200, 255, 255)
    return f


def mixed(width: int, height: int, index: int) -> np.ndarray:
    """Frame ``index`` of BASELINE config 5's batch: seeds 0..63, cycling
    random / ramp / solid / natural."""
    k = index % 4
    if k == 0:
        return random(width, height, seed=index)
    if k == 1:
        return ramp(width, height)
    if k == 2:
        g = (37 * index + 11) & 255
        return solid(width, height, (g, (g * 3) & 255, (g * 7) & 255, 255))
    return natural(width, height, seed=index)


def with_pitch(frame: np.ndarray, linesize: int, fill: int = 0xA5) -> np.ndarray:
    """Copy an (H, W, 4) frame into an (H, linesize) buffer whose padding bytes
    hold garbage, like a driver-pitched staging surface."""
    h, w, _ = frame.shape
    assert linesize >= w * 4
    buf = np.full((h, linesize), fill, np.uint8)
    buf[:, : w * 4] = frame.reshape(h, w * 4)
    return buf
