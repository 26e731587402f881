"""Quality / compression metrics of the reference's analysis script, without matplotlib.

Definitions follow analyze_results.py of the reference repository: both images are opened with
PIL and converted to 'L' (:57-63), MSE over float64 pixels (:17-24), PSNR = 20*log10(255/sqrt(MSE))
(:26-32), compression ratio = original bytes / compressed bytes (:66-72), bits per pixel =
compressed bytes * 8 / pixels (:74-76).  SSIM (:83-84) is reported only when scikit-image is
installed.  Host-side reporting helper; nothing here is on the encode path.
"""
from __future__ import annotations

import math
import os

import numpy as np


def calculate_mse(image1, image2) -> float:
    a = np.asarray(image1, dtype=np.float64)
    b = np.asarray(image2, dtype=np.float64)
    return float(np.mean((a - b) ** 2))


def calculate_psnr(mse: float) -> float:
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def analyze(original_path: str, compressed_path: str) -> dict:
    from PIL import Image
    orig = Image.open(original_path).convert("L")
    comp = Image.open(compressed_path).convert("L")
    if orig.size != comp.size:
        orig = orig.resize(comp.size)
    size_o, size_c = os.path.getsize(original_path), os.path.getsize(compressed_path)
    w, h = orig.size
    mse = calculate_mse(orig, comp)
    out = {"file_size_orig": size_o, "file_size_comp": size_c,
           "compression_ratio": size_o / size_c if size_c else 0.0, "bpp": size_c * 8 / (w * h),
           "mse": mse, "psnr": calculate_psnr(mse), "ssim": None}
    try:
        from skimage.metrics import structural_similarity
        out["ssim"] = float(structural_similarity(np.asarray(orig), np.asarray(comp), data_range=255))
    except ImportError:
        pass
    return out


def coefficient_mismatches(ours: np.ndarray, reference: np.ndarray) -> dict:
    """north_star: any +-1 coefficient mismatch from DCT rounding is counted and reported."""
    d = ours.astype(np.int32) - reference.astype(np.int32)
    return {"total": int(d.size), "mismatched": int(np.count_nonzero(d)), "off_by_one": int(np.count_nonzero(np.abs(d) == 1)),
            "max_abs": int(np.abs(d).max()) if d.size else 0}
