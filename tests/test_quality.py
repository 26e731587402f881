"""Metric definitions of the reference's analyze_results.py (MSE / PSNR / CR / bpp), restated."""
import math
import os

import numpy as np
import pytest

from jpeg_image_compression_b200.quality import analyze, calculate_mse, calculate_psnr, coefficient_mismatches


def test_mse_psnr_definitions():
    a = np.zeros((4, 4), np.uint8)
    b = np.full((4, 4), 5, np.uint8)
    assert calculate_mse(a, b) == 25.0 and calculate_mse(a, a) == 0.0
    assert calculate_psnr(0) == float("inf")
    assert math.isclose(calculate_psnr(25.0), 20 * math.log10(255 / 5))
    m = coefficient_mismatches(np.array([1, 2, 3, 5]), np.array([1, 3, 3, 2]))
    assert m == {"total": 4, "mismatched": 2, "off_by_one": 1, "max_abs": 3}


def test_analyze_golden_file(golden, tmp_path):
    """Decode a reference-built JPEG with PIL: the numbers are those the reference's script would print."""
    from oracle.oracle import write_bmp
    rgb = golden["lena_crop256/rgb"]
    bmp, jpg = str(tmp_path / "a.bmp"), str(tmp_path / "a.jpg")
    write_bmp(bmp, rgb)
    open(jpg, "wb").write(golden["lena_crop256/file"].tobytes())
    q = analyze(bmp, jpg)
    assert q["file_size_comp"] == golden["lena_crop256/file"].size
    assert math.isclose(q["bpp"], q["file_size_comp"] * 8 / (256 * 256))
    assert math.isclose(q["compression_ratio"], os.path.getsize(bmp) / q["file_size_comp"])
    assert 30.0 < q["psnr"] < 45.0 and math.isclose(q["psnr"], calculate_psnr(q["mse"]))


@pytest.mark.skipif(not os.path.isdir("/root/reference/assets/input"), reason="reference assets only in the build container")
def test_psnr_matches_survey_numbers(ref, tmp_path):
    """natural_c output of lena decodes to 35.76 dB, blackbuck 42.50 dB."""
    import ctypes as C
    from oracle.oracle import _Img
    for name, want in (("lena", 35.76), ("blackbuck", 42.50)):
        bmp = f"/root/reference/assets/input/{name}.bmp"
        rgb = ref.load_bmp(bmp)
        scan = ref.encode_scan(rgb)
        from oracle.oracle import Oracle
        jpg = str(tmp_path / f"{name}.jpg")
        open(jpg, "wb").write(Oracle().jfif_header(rgb.shape[1], rgb.shape[0]) + scan + b"\xff\xd9")
        assert abs(analyze(bmp, jpg)["psnr"] - want) < 0.01, name
