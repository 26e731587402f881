"""CPU tests for the host side: the C-ABI library loads and exports every symbol that
include/jpegb200.h declares (no compute calls: there is no GPU here), the host-side BMP reader
and JFIF header writer match the reference, and the synthetic generators agree."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import jpeg_image_compression_b200 as jb
from jpeg_image_compression_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    jb.build_library()
    return jb.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "jpegb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(\w+)\s*\(", header)) - {"defined", "sizeof"}
    declared = {d for d in declared if not d.isupper()}
    assert len(declared) >= 36
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/jpegb200.h but not exported"
    for name in _lib.EXPORTED_FUNCTIONS:
        assert name in declared, f"{name} bound in python but not declared in the header"
    for name in _lib.EXPORTED_DATA:
        C.c_ubyte.in_dll(lib, name)


def test_struct_layouts_match_reference():
    # sizes verified against the reference with sizeof on x86-64
    assert C.sizeof(_lib.BMPImage) == 16 and C.sizeof(_lib.PlaneImage) == 16
    assert C.sizeof(_lib.ZigZagData) == 24 and C.sizeof(_lib.RLEData) == 24 and C.sizeof(_lib.JpegEncoderBuffer) == 24
    from jpeg_image_compression_b200.stages import SYMBOL_DTYPE
    assert SYMBOL_DTYPE.itemsize == 6


def test_tables_identical_to_oracle(lib, oracle):
    for name, oname, n in (("std_luminance_quant_tbl", "orc_quant_luma", 64), ("std_dc_luminance_nrcodes", "orc_dc_counts", 16),
                           ("std_dc_luminance_values", "orc_dc_values", 12), ("std_ac_luminance_nrcodes", "orc_ac_counts", 16),
                           ("std_ac_luminance_values", "orc_ac_values", 162)):
        ours = bytes((C.c_ubyte * n).in_dll(lib, name))
        theirs = bytes((C.c_ubyte * n).in_dll(oracle.lib, oname))
        assert ours == theirs, name


def test_tables_identical_to_reference(lib, ref):
    for name, n in (("std_luminance_quant_tbl", 64), ("std_dc_luminance_nrcodes", 16), ("std_dc_luminance_values", 12),
                    ("std_ac_luminance_nrcodes", 16), ("std_ac_luminance_values", 162)):
        assert bytes((C.c_ubyte * n).in_dll(lib, name)) == bytes((C.c_ubyte * n).in_dll(ref.lib, name)), name


def test_bmp_reader_matches_reference_goldens(lib, golden):
    for name in golden["bmp_names"]:
        got = jb.loadBMPImage(os.path.join(ROOT, "tests", "golden", "bmp", f"{name}.bmp"))
        assert got is not None and np.array_equal(got, golden[f"bmp/{name}"]), name


def test_bmp_reader_rejects_like_reference(lib, tmp_path, capfd):
    bad = tmp_path / "bad.bmp"
    bad.write_bytes(b"PNG" + bytes(100))
    assert jb.loadBMPImage(str(bad)) is None
    assert "not a valid BMP" in capfd.readouterr().err
    assert jb.loadBMPImage(str(tmp_path / "missing.bmp")) is None
    from oracle.oracle import write_bmp
    p = tmp_path / "p8.bmp"
    write_bmp(str(p), np.zeros((2, 2, 3), np.uint8))
    raw = bytearray(p.read_bytes())
    raw[28] = 8                                                  # biBitCount = 8
    p.write_bytes(bytes(raw))
    assert jb.loadBMPImage(str(p)) is None
    assert "Only 24-bit" in capfd.readouterr().err
    raw[28] = 24
    raw[30] = 1                                                  # biCompression = 1
    p.write_bytes(bytes(raw))
    assert jb.loadBMPImage(str(p)) is None
    raw[30] = 0
    p.write_bytes(bytes(raw[:-3]))                               # truncated pixel data
    assert jb.loadBMPImage(str(p)) is None
    assert "Insufficient data" in capfd.readouterr().err


def test_bmp_reader_vs_reference_random(lib, ref, tmp_path):
    from oracle.oracle import write_bmp
    rng = np.random.default_rng(3)
    for i, (w, h, td, hs) in enumerate([(1, 1, False, 40), (5, 7, True, 40), (33, 9, False, 124), (64, 3, True, 108)]):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / f"r{i}.bmp")
        write_bmp(p, rgb, top_down=td, header_size=hs)
        a, b = jb.loadBMPImage(p), ref.load_bmp(p)
        assert np.array_equal(a, b) and np.array_equal(a, rgb)


def test_jfif_header_matches_goldens(lib, golden, golden_names):
    for name in golden_names:
        h, w, _ = golden[f"{name}/rgb"].shape
        assert jb.jfif_header(w, h) == golden[f"{name}/file"].tobytes()[:328], name


def test_synth_numpy_equals_oracle(oracle):
    for (w, h, seed, amp) in [(64, 48, 1, 20), (131, 77, 12345, 0), (300, 5, 7, 64), (1920, 16, 4095, 20)]:
        assert np.array_equal(jb.synth_rgb(w, h, seed, amp), oracle.synth_rgb(w, h, seed, amp))


def test_cli_usage_and_load_failure(lib, tmp_path):
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_app")
    r = subprocess.run([app], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr and "<input_file_path> <output_file_path>" in r.stderr
    r = subprocess.run([app, str(tmp_path / "nope.bmp"), str(tmp_path / "o.jpg")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error: Failed to load image from" in r.stderr
    assert r.stdout.startswith("Starting processing...\nInput: ")


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.jpegb200_device_count() == 0
    assert not lib.jpegb200_encoder_create(0)
    with pytest.raises(jb.JpegB200Error):
        jb.DeviceEncoder(0)
    with pytest.raises(jb.JpegB200Error):
        jb.encode_scan_host(np.zeros((8, 8, 3), np.uint8))
    assert jb.convertBMPToJPEGGrayscale(np.zeros((8, 8, 3), np.uint8)) is None


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jpeg_image_compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".c", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_batch_cli_usage(lib):
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_batch")
    r = subprocess.run([app], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
