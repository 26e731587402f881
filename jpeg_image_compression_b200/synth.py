"""Synthetic workload generator, numpy form.

Integer-only and counter-based, so it produces exactly the bytes of the device generator
(csrc/synth.cuh) and of the oracle's C version; used to build host-side inputs."""
from __future__ import annotations

import numpy as np


def _tri(t: np.ndarray, period: int) -> np.ndarray:
    m = t % np.uint32(2 * period)
    d = np.where(m > period, m - np.uint32(period), np.uint32(period) - m)
    return (d * np.uint32(255)) // np.uint32(period)


def synth_rgb(w: int, h: int, seed: int = 1, amp: int = 20) -> np.ndarray:
    x = np.arange(w, dtype=np.uint32)[None, :]
    y = np.arange(h, dtype=np.uint32)[:, None]
    with np.errstate(over="ignore"):
        base = (_tri(x + np.uint32(2) * y, 419) + _tri(np.uint32(3) * x + np.uint32(1 << 20) - y, 1021)
                + _tri(y + np.uint32(0) * x, 173) + _tri(x + np.uint32(0) * y, 67)) // np.uint32(4)
        out = np.empty((h, w, 3), np.uint8)
        span = np.uint32(2 * amp + 1)
        for c, off in enumerate((10, 0, -10)):
            k = np.uint32((c * 0xC2B2AE3D) & 0xFFFFFFFF) ^ np.uint32((seed * 0x27D4EB2F) & 0xFFFFFFFF)
            hsh = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ k
            hsh ^= hsh >> np.uint32(15)
            hsh *= np.uint32(0x2C1B3C6D)
            hsh ^= hsh >> np.uint32(12)
            hsh *= np.uint32(0x297A2D39)
            hsh ^= hsh >> np.uint32(15)
            noise = ((hsh >> np.uint32(24)) % span).astype(np.int32) - amp
            out[:, :, c] = np.clip(base.astype(np.int32) + off + noise, 0, 255).astype(np.uint8)
    return out
