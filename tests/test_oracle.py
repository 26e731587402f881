"""CPU tests: pin the oracle (oracle/jpeg_oracle.c) to the reference.

1. against the committed golden fixtures minted from the reference's own objects
   (tests/golden/make_golden.py), every stage, bit for bit;
2. against oracle/_ref (the unmodified reference compiled here) on fresh random and
   synthetic inputs and -- when /root/reference is present -- on its four asset BMPs.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle.oracle import SYMBOL_DTYPE, pad8

ASSETS = "/root/reference/assets/input"


def _sym_fields(sym):
    return np.stack([sym["symbol"].astype(np.uint16), sym["amplitude"], sym["nbits"].astype(np.uint16)], 1)


def _golden_sym_fields(g):          # stored as bytes [symbol, amp_lo, amp_hi, nbits]
    g = g.astype(np.uint16)
    return np.stack([g[:, 0], g[:, 1] | (g[:, 2] << 8), g[:, 3]], 1)


def test_golden_every_stage(oracle, golden, golden_names):
    assert len(golden_names) >= 15
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        y = oracle.luma_pad(rgb)
        assert np.array_equal(y, golden[f"{name}/y"]), name
        d = oracle.fdct_image(oracle.level_shift(y))
        assert np.array_equal(d.view(np.uint32), golden[f"{name}/dct"].view(np.uint32)), name
        zz = oracle.zigzag(oracle.quantize(d))
        assert np.array_equal(zz, golden[f"{name}/zigzag"]), name
        sym = oracle.rle(zz)
        assert np.array_equal(_sym_fields(sym), _golden_sym_fields(golden[f"{name}/symbols"])), name
        scan = oracle.huffman(sym, zz.shape[0])
        assert scan.tobytes() == golden[f"{name}/scan"].tobytes(), name
        assert oracle.encode_scan(rgb) == golden[f"{name}/scan"].tobytes(), name
        assert oracle.encode_file_bytes(rgb) == golden[f"{name}/file"].tobytes(), name
        assert np.array_equal(oracle.coefficients(rgb), golden[f"{name}/zigzag"]), name


def test_block_bits_sum_matches_scan(oracle, golden, golden_names):
    for name in golden_names:
        zz = golden[f"{name}/zigzag"]
        bits = int(oracle.block_bits(zz).sum())
        scan = golden[f"{name}/scan"]
        stuffed = int((scan == 0xFF).sum())            # every 0xFF in a stuffed scan is followed by 0x00
        assert (bits + 7) // 8 == scan.size - stuffed, name


def test_header_is_328_bytes_and_carries_unpadded_dims(oracle):
    h = oracle.jfif_header(762, 1309)
    assert len(h) == 328 and h[:4] == b"\xff\xd8\xff\xe0"
    assert h[94:98] == (1309).to_bytes(2, "big") + (762).to_bytes(2, "big")
    assert h[-10:-8] == b"\xff\xda"


def test_synth_generator_hashes(oracle, synth_hashes):
    e = synth_hashes["1920x1080_seed0_amp20"]
    rgb = oracle.synth_rgb(1920, 1080, 0, 20)
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == e["rgb_sha256"]


def test_full_size_synthetic_vs_reference_hash(oracle, synth_hashes):
    e = synth_hashes["1920x1080_seed64_amp20"]
    rgb = oracle.synth_rgb(e["w"], e["h"], e["seed"], e["amp"])
    zz = oracle.coefficients(rgb)
    assert hashlib.sha256(zz.tobytes()).hexdigest() == e["zigzag_sha256"]
    scan = oracle.encode_scan(rgb)
    assert len(scan) == e["scan_bytes"]
    assert hashlib.sha256(scan).hexdigest() == e["scan_sha256"]


def test_against_reference_objects_random(oracle, ref):
    rng = np.random.default_rng(99)
    for (w, h) in [(1, 1), (3, 2), (8, 8), (9, 8), (16, 15), (31, 33), (100, 37), (257, 64)]:
        for kind in ("noise", "smooth"):
            rgb = (rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if kind == "noise"
                   else oracle.synth_rgb(w, h, int(rng.integers(1 << 30)), int(rng.integers(0, 40))))
            st = ref.stages(rgb)
            y = oracle.luma_pad(rgb)
            assert y.shape == (pad8(h), pad8(w)) and np.array_equal(y, st["y"])
            c = oracle.level_shift(y)
            assert np.array_equal(c, st["centered"])
            d = oracle.fdct_image(c)
            assert np.array_equal(d.view(np.uint32), st["dct"].view(np.uint32))
            q = oracle.quantize(d)
            assert np.array_equal(q, st["quant"])
            zz = oracle.zigzag(q)
            assert np.array_equal(zz, st["zigzag"])
            sym = oracle.rle(zz)
            assert np.array_equal(_sym_fields(sym), _sym_fields(st["symbols"]))
            assert oracle.huffman(sym, zz.shape[0]).tobytes() == st["scan"]
            assert oracle.encode_scan(rgb) == st["scan"]


def test_dct_block_bitwise_vs_reference(oracle, ref):
    rng = np.random.default_rng(5)
    for _ in range(200):
        blk = rng.integers(-128, 128, (8, 8), dtype=np.int8)
        assert np.array_equal(oracle.fdct_block(blk).view(np.uint32), ref.fdct_block(blk).view(np.uint32))
    for v in (-128, 127, 0):
        blk = np.full((8, 8), v, np.int8)
        assert np.array_equal(oracle.fdct_block(blk).view(np.uint32), ref.fdct_block(blk).view(np.uint32))


@pytest.mark.skipif(not os.path.isdir(ASSETS), reason="reference assets only exist in the build container")
def test_reference_assets(oracle, ref, synth_hashes):
    for name in ["lena", "blackbuck", "greenland", "offset_sample"]:
        rgb = ref.load_bmp(f"{ASSETS}/{name}.bmp")
        e = synth_hashes[f"asset_{name}"]
        assert hashlib.sha256(rgb.tobytes()).hexdigest() == e["rgb_sha256"]
        scan = oracle.encode_scan(rgb)
        assert scan == ref.encode_scan(rgb)
        assert hashlib.sha256(scan).hexdigest() == e["scan_sha256"] and len(scan) == e["scan_bytes"]
        assert hashlib.sha256(oracle.coefficients(rgb).tobytes()).hexdigest() == e["zigzag_sha256"]


def test_quantized_magnitude_bound(oracle):
    """|q| <= 95 for any input (the device stores coefficients as int8): check the
    analytic extreme blocks (sign patterns of each basis function at full swing)."""
    worst = 0
    for u in range(8):
        for v in range(8):
            r = np.arange(8)
            bu = np.cos((2 * r + 1) * u * np.pi / 16)
            bv = np.cos((2 * r + 1) * v * np.pi / 16)
            pat = np.sign(np.outer(bu, bv))
            for s in (1, -1):
                blk = np.where(s * pat >= 0, 127, -128).astype(np.int8)
                q = oracle.quantize(oracle.fdct_block(blk).reshape(8, 8))
                worst = max(worst, int(np.abs(q).max()))
    assert worst <= 95


def test_banded_streaming_oracle_equals_whole_image_oracle(oracle, synth_hashes):
    """orc_encode_scan_synth_banded (the oracle for images beyond the reference's int-index limit,
    SURVEY.md 8c) walks the image in bands with carried entropy state: it must equal the whole-image
    oracle for every band size / thread count, and the reference build's hashes at 4K."""
    for (w, h, seed, amp, rows, threads) in ((200, 120, 3, 20, 1, 1), (203, 77, 5, 30, 2, 3), (64, 8, 1, 0, 8, 4),
                                            (1283, 725, 9, 20, 4, 8), (7, 9, 2, 64, 1, 2), (257, 1031, 4, 64, 3, 5)):
        whole = oracle.encode_scan(oracle.synth_rgb(w, h, seed, amp))
        assert oracle.encode_scan_synth_banded(w, h, seed, amp, rows, threads) == whole, (w, h, rows, threads)
    e = synth_hashes["3840x2160_seed1_amp20"]
    b = oracle.encode_scan_synth_banded(3840, 2160, 1, 20, 8, 0)
    assert len(b) == e["scan_bytes"] and hashlib.sha256(b).hexdigest() == e["scan_sha256"]
