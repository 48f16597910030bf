"""Synthetic frames used by every parity test and by bench.py.  Pure numpy, no reference code."""
import numpy as np


def _lcg_states(seed, n):
    """s_k = a*s_{k-1} + c mod 2^32 for k=1..n, vectorised: s_k = a^k*s0 + c*(a^(k-1)+...+1)."""
    a = np.uint32(1664525)
    c = np.uint32(1013904223)
    with np.errstate(over="ignore"):
        pw = np.empty(n + 1, dtype=np.uint32)
        pw[0] = 1
        pw[1:] = a
        pw = np.cumprod(pw, dtype=np.uint32)             # a^0 .. a^n   (mod 2^32)
        geo = np.cumsum(pw[:-1], dtype=np.uint32)        # sum_{j<k} a^j for k=1..n
        return pw[1:] * np.uint32(seed & 0xffffffff) + c * geo


def frame_g(width, height, seed=12345, stride=None):
    """G(seed): gradient + 97x61 checker + dark bars + 3-bit LCG noise (the frame of SURVEY section 6/8d)."""
    stride = stride or width
    s = _lcg_states(seed, width * height).reshape(height, width)
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    v = 96 + (x * 37) // width + np.where(((x // 97 + y // 61) & 1) == 1, 70, 0) \
        + np.where((x % 211 < 9) & (y % 173 < 60), -80, 0) + ((s >> np.uint32(24)) & np.uint32(7)).astype(np.int64)
    out = np.zeros((height, stride), np.uint8)
    out[:, :width] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def frame_uniform(width, height, seed=1, stride=None):
    """U: uniform random bytes (worst case for hysteresis / FAST early-out / SHT)."""
    stride = stride or width
    rng = np.random.default_rng(seed)
    out = np.zeros((height, stride), np.uint8)
    out[:, :width] = rng.integers(0, 256, (height, width), dtype=np.uint8)
    return out


def frame_text(width, height, seed=7, stride=None):
    """T: white background with black 5x9 glyph-like rectangles on a 12x16 grid (threshold / CCL / MSER)."""
    stride = stride or width
    rng = np.random.default_rng(seed)
    img = np.full((height, stride), 255, np.uint8)
    for gy in range(4, height - 12, 16):
        for gx in range(4, width - 8, 12):
            if rng.random() < 0.8:
                h = int(rng.integers(5, 10))
                w = int(rng.integers(3, 6))
                img[gy:gy + h, gx:gx + w] = int(rng.integers(0, 60))
                if rng.random() < 0.3:  # a hole, so that regions are not all convex
                    img[gy + 2:gy + h - 2, gx + 1:gx + w - 1] = 255
    img[:, width:] = 0
    return img


def frame_smooth(width, height, seed=3, stride=None):
    """Smooth blobs + lines: long connected edges (exercises cross-tile hysteresis and KHT linking)."""
    stride = stride or width
    rng = np.random.default_rng(seed)
    x = np.arange(width, dtype=np.float32)[None, :]
    y = np.arange(height, dtype=np.float32)[:, None]
    v = np.full((height, width), 100.0, np.float32)
    for _ in range(6):
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        r = rng.uniform(min(width, height) / 10, min(width, height) / 3)
        v += 60.0 * (((x - cx) ** 2 + (y - cy) ** 2) < r * r)
    for _ in range(5):
        a, b = rng.uniform(-1, 1), rng.uniform(0, height)
        v += 50.0 * (np.abs(y - (a * x + b)) < 2.5)
    v += rng.normal(0, 1.5, (height, width)).astype(np.float32)
    out = np.zeros((height, stride), np.uint8)
    out[:, :width] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def frame_const(width, height, value=0, stride=None):
    stride = stride or width
    out = np.zeros((height, stride), np.uint8)
    out[:, :width] = value
    return out
