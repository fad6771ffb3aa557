"""Device-side generators of the synthetic BGRA batches the benchmark streams (torch owns the
memory).  Same families as frames.py (random / ramp / solid / natural, cycling like
BASELINE config 5's 64-frame batch); the random values come from torch's generator, so
parity tests use frames.py + the oracle, and bench.py uses these for bulk data."""
from __future__ import annotations

import torch


def _ramp(w, h, dev):
    x = torch.arange(w, device=dev, dtype=torch.int32)[None, :]
    y = torch.arange(h, device=dev, dtype=torch.int32)[:, None]
    f = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    f[..., 0] = (x & 255).expand(h, w).to(torch.uint8)
    f[..., 1] = (y & 255).expand(h, w).to(torch.uint8)
    f[..., 2] = ((x + y) & 255).to(torch.uint8)
    f[..., 3] = 255
    return f


def _random(w, h, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f = torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device=dev, generator=g)
    f[..., 3] = 255
    return f


def _solid(w, h, dev, index):
    g = (37 * index + 11) & 255
    f = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    f[..., 0], f[..., 1], f[..., 2], f[..., 3] = g, (g * 3) & 255, (g * 7) & 255, 255
    return f


def _natural(w, h, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + seed)
    x = torch.linspace(0, 1, w, device=dev)[None, :]
    y = torch.linspace(0, 1, h, device=dev)[:, None]
    base = [96 + 80 * torch.sin(2.1 * x + 0.7 * y + seed), 110 + 70 * torch.cos(1.3 * x - 1.9 * y + 0.3 * seed),
            128 + 60 * torch.sin(0.9 * x * y * 3.0 + 1.1)]
    f = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    for c in range(3):
        n = torch.randint(-3, 4, (h, w), device=dev, generator=g)
        f[..., c] = (base[c] + n).clamp(0, 255).to(torch.uint8)
    f[..., 3] = 255
    return f


def _ui(w, h, dev, seed):
    """frames.ui on the device (integer arithmetic only: bit-identical to the numpy version)"""
    x = torch.arange(w, device=dev, dtype=torch.int64)[None, :]
    y = torch.arange(h, device=dev, dtype=torch.int64)[:, None]
    f = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    f[..., 0], f[..., 1], f[..., 2], f[..., 3] = 34, 30, 30, 255
    line = (y % 24 >= 6) & (y % 24 < 18)
    hsh = (((x // 2) * 73856093) ^ ((y // 2) * 19349663) ^ (seed * 83492791)) & 0xFFFFFFFF
    glyph = line & (((hsh >> 7) & 3) == 0) & (x % 512 < 400)
    word = ((((x // 64) * 2654435761) ^ ((y // 24) * 40503) ^ seed) & 0xFFFFFFFF) >> 11
    accent = (word % 5 == 0) & glyph
    plain = glyph & ~accent
    f[plain] = torch.tensor([220, 220, 220, 255], dtype=torch.uint8, device=dev)
    f[accent] = torch.tensor([90, 200, 255, 255], dtype=torch.uint8, device=dev)
    return f


def mixed_frame(w: int, h: int, index: int, dev) -> torch.Tensor:
    k = index % 4
    if k == 0:
        return _random(w, h, dev, index)
    if k == 1:
        return _ramp(w, h, dev)
    if k == 2:
        return _solid(w, h, dev, index)
    return _natural(w, h, dev, index)


def mixed_batch(n: int, w: int, h: int, dev, first_index: int = 0, content: str = "mixed") -> torch.Tensor:
    """(n, h, w, 4) uint8 on `dev`; frame i has global index first_index + i."""
    out = torch.empty((n, h, w, 4), dtype=torch.uint8, device=dev)
    for i in range(n):
        idx = first_index + i
        if content == "mixed":
            out[i] = mixed_frame(w, h, idx, dev)
        elif content == "random":
            out[i] = _random(w, h, dev, idx)
        elif content == "natural":
            out[i] = _natural(w, h, dev, idx)
        elif content == "ramp":
            out[i] = _ramp(w, h, dev)
        elif content == "solid":
            out[i] = _solid(w, h, dev, idx)
        elif content == "ui":
            out[i] = _ui(w, h, dev, idx)
        else:
            raise ValueError(content)
    return out
